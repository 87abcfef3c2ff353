#!/bin/bash
# tools/run_with_deadline.sh SECONDS cmd...: run cmd in its own process group; at the deadline send SIGABRT to the whole
# group (PYTHONFAULTHANDLER=1 makes every python process print its threads' tracebacks), then SIGKILL.
d=$1; shift
export PYTHONFAULTHANDLER=1
setsid "$@" &
pid=$!
( sleep "$d"; echo "[deadline] $d s: aborting process group $pid"; kill -ABRT -- -$pid 2>/dev/null; sleep 4; kill -KILL -- -$pid 2>/dev/null ) &
w=$!
wait $pid; rc=$?
kill $w 2>/dev/null
exit $rc
