"""CPU: marllb_b200/wire.py against bytes produced by the unmodified reference layouts
(tests/golden/make_wire_golden.py -> wire.npz): msg_out / msg_in records byte for byte, the alias
table of rl_controller._build_alias_table exactly, and the round trips."""
import numpy as np

from conftest import load_golden


def test_sizes_match_reference_layouts():
    from marllb_b200 import wire
    g = load_golden("wire")
    assert [wire.MSG_OUT_SIZE, wire.MSG_IN_SIZE] == g["sizes"][:2].tolist() == [2853, 792]
    assert wire.MSG_OUT.fields["as_stats"][1] == int(g["sizes"][2]) and wire._SERVER.itemsize == int(g["sizes"][3])


def test_msg_out_bytes_and_round_trip():
    from marllb_b200 import wire
    g = load_golden("wire")
    for k in range(int(g["out_n"])):
        obs = g[f"out{k}_obs"]
        rec = wire.obs_to_msg_out(obs, sequence_id=1000 + k, timestamp_us=123456789 + k)
        assert rec.tobytes() == g[f"out{k}_bytes"].tobytes(), k
        back = wire.msg_out_to_obs(g[f"out{k}_bytes"].tobytes(), obs.shape[0])[0]
        want = obs.copy()
        want[:, 0] = np.floor(want[:, 0])                       # n_flow_on travels as uint32
        assert np.array_equal(back, want)
        active = np.flatnonzero((obs > 0).any(axis=1))
        assert np.array_equal(active, g[f"out{k}_active"])
    # batched: E records at once equal E single records
    obs = np.stack([g["out4_obs"], g["out5_obs"]])
    recs = wire.obs_to_msg_out(obs, sequence_id=5, timestamp_us=9)
    assert recs[0].tobytes() == wire.obs_to_msg_out(obs[0], 5, 9).tobytes()
    assert recs[1].tobytes() == wire.obs_to_msg_out(obs[1], 5, 9).tobytes()


def test_alias_table_and_msg_in_bytes():
    from marllb_b200 import wire
    g = load_golden("wire")
    for k in range(int(g["in_n"])):
        w = g[f"in{k}_w"]
        prob, alias = wire.build_alias_table(w)
        assert np.array_equal(prob, g[f"in{k}_prob"]) and np.array_equal(alias, g[f"in{k}_alias"])
        rec = wire.actions_to_msg_in(w, sequence_id=77 + k, timestamp_us=int(1700000000.25 * 1e6))
        assert rec.tobytes() == g[f"in{k}_bytes"].tobytes(), k
        un = wire.msg_in_to_actions(rec.tobytes())[0]
        assert un["num_servers"] == len(w) and np.allclose(un["weights"], w, rtol=1e-6)
        # the table samples the weights: exact expectation of the alias method
        n = len(w)
        p = np.zeros(n)
        for b in range(n):
            p[b] += prob[b] / n
            p[alias[b]] += (1.0 - prob[b]) / n
        assert np.allclose(p, w, atol=1e-12)


def test_weights_from_server_action():
    from marllb_b200 import wire
    w = wire.weights_from_server_action(np.array([[0, 1, 2, 1], [2, 2, 2, 2]]))
    assert np.allclose(w[0], np.array([1.0, 1.5, 2.0, 1.5]) / 6.0) and np.allclose(w[1], 0.25)
