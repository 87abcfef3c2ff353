"""CPU: the full-size policy fixture and the input spec it was generated from stay in step (tests/fullsize_spec.py is
imported by both the generator, which runs the reference in the build container, and the GPU test)."""
import numpy as np

import fullsize_spec as spec
from conftest import load_golden


def test_spec_is_deterministic_and_sized_like_the_benchmarked_configs():
    assert spec.SAC["state_dim"] == 256 * 11 and spec.SAC["batch_size"] == 256          # BASELINE config 4
    assert spec.QMIX["num_agents"] == 2 and spec.QMIX["obs_dim"] == 32 * 11             # BASELINE config 3
    a, b = spec.sac_batch(1), spec.sac_batch(1)
    assert all(np.array_equal(x, y) for x, y in zip(a[0], b[0])) and np.array_equal(a[1], b[1])
    assert not np.array_equal(spec.sac_batch(1)[0][0], spec.sac_batch(2)[0][0])
    q = spec.qmix_batch(1)
    L = q["seq_lengths"]
    assert L.max() == spec.QMIX["max_seq_len"] and L.min() >= spec.QMIX["max_seq_len"] // 2
    assert all(q["dones"][i, l - 1] == 1.0 and not q["observations"][i, l:].any() for i, l in enumerate(L))
    sd = spec.synth_state_dict({"w": (8, 4), "b": (8,)}, 3)
    assert sd["w"].dtype == np.float32 and np.array_equal(sd["w"], spec.synth_state_dict({"b": (8,), "w": (8, 4)}, 3)["w"])


def test_fixture_holds_every_update_of_both_learners():
    g = load_golden("policy_fullsize")
    for u in range(1, spec.N_UPDATES + 1):
        assert g[f"sac.upd{u}.losses"].shape == (4,) and np.isfinite(g[f"sac.upd{u}.losses"]).all()
        assert g[f"qmix.upd{u}.stats"].shape == (3,) and np.isfinite(g[f"qmix.upd{u}.stats"]).all()
    name = "sac.m1.q1.gru.weight_ih_l0"
    assert g[name].shape == (spec.N_SAMPLES,) and float(g[name + ".absmax"][0]) > 0
    idx = spec.sample_index("gru.weight_ih_l0", 384 * (256 * 11 + 256), 99)
    assert idx.shape == (spec.N_SAMPLES,) and idx.max() < 384 * (256 * 11 + 256)
    # the sampled gradient entries are not all zero: the fixture really carries the reference's first Adam moments
    assert np.count_nonzero(g[name]) > spec.N_SAMPLES // 2
