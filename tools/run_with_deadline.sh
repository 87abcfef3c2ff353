#!/bin/bash
# tools/run_with_deadline.sh SECONDS cmd...: run cmd in its own process group; at the deadline send SIGABRT to the whole
# group (PYTHONFAULTHANDLER=1 makes every python process print its threads' tracebacks), then SIGKILL.
# The watchdog is a process group of its own with its output in a file, so that it can be removed as a whole and a
# sleeping child never keeps a `| tail` on our stdout waiting.
d=$1; shift
export PYTHONFAULTHANDLER=1
setsid "$@" &
pid=$!
log=/tmp/deadline_$$.log
setsid bash -c "sleep $d; echo '[deadline] $d s: aborting process group $pid'; kill -ABRT -- -$pid 2>/dev/null; sleep 4; kill -KILL -- -$pid 2>/dev/null" > $log 2>&1 &
w=$!
wait $pid; rc=$?
kill -KILL -- -$w 2>/dev/null
[ -s $log ] && cat $log >&2
exit $rc
