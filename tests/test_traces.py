"""CPU: trace parsing and chunking.  No GPU needed: TraceStream is host logic."""
import numpy as np

from marllb_b200 import traces


def _write_tsv(path, t, n):
    with open(path, "w") as f:
        f.write("time\tquery\n")
        for a, b in zip(t, n):
            f.write(f"{a:.6f}\t/dummy.php/?n={int(b)}\n")


def test_trace_stream_chunks_partition_the_file(tmp_path):
    rng = np.random.RandomState(0)
    t = np.cumsum(rng.exponential(1 / 40.0, 900))
    n = rng.randint(11, 38_400_000, len(t))
    p = str(tmp_path / "rate_40.csv")
    _write_tsv(p, t, n)
    whole = traces.load_trace(p)                                   # the one-shot loader
    A, dt, spc = 3, 0.25, 5
    ts = traces.TraceStream(p, num_envs=2, num_agents=A, dt=dt, steps_per_chunk=spc)
    per_stream = [[] for _ in range(2 * A)]
    per_work = [[] for _ in range(2 * A)]
    k_prev = 0
    while not ts.exhausted:
        k_end, ch = ts.next_chunk()
        assert k_end == k_prev + spc
        lo, hi = np.float32(np.float32(k_prev) * np.float32(dt)), np.float32(np.float32(k_end) * np.float32(dt))
        off = ch["offsets"]
        assert off.shape == (2 * A + 1,) and off[-1] == len(ch["time"])
        for s in range(2 * A):
            seg = ch["time"][off[s]:off[s + 1]]
            assert seg.dtype == np.float32 and np.all(seg < hi) and np.all(seg >= lo)   # exactly this chunk's windows
            per_stream[s].append(seg)
            per_work[s].append(ch["work"][off[s]:off[s + 1]])
        k_prev = k_end
    ref = traces.split_round_robin(whole, A)                       # row r -> agent r mod A
    for e in range(2):                                             # both envs replay the same file
        for a in range(A):
            assert np.array_equal(np.concatenate(per_stream[e * A + a]), ref[a]["time"])
            assert np.array_equal(np.concatenate(per_work[e * A + a]), ref[a]["work"])
    ts.restart()
    assert ts.next_chunk()[0] == spc


def test_trace_stream_dict_sources_with_alias_randoms():
    rng = np.random.RandomState(1)
    srcs = [traces.poisson_trace(30.0, 6.0, 0.1, rng, servers=4) for _ in range(3)]
    ts = traces.TraceStream(srcs, num_envs=3, num_agents=2, steps_per_chunk=4)
    got_b = [[] for _ in range(6)]
    while not ts.exhausted:
        _, ch = ts.next_chunk()
        off = ch["offsets"]
        if off[-1]:
            assert ch["bucket"].dtype == np.int32 and ch["u"].dtype == np.float32
            for s in range(6):
                got_b[s].append(ch["bucket"][off[s]:off[s + 1]])
    for e in range(3):
        for a in range(2):
            assert np.array_equal(np.concatenate(got_b[e * 2 + a]), srcs[e]["bucket"][a::2])
