"""2 GPUs, NCCL: a QMIX update over a global batch on one rank equals the same update with the
batch split over two ranks (flat-bucket gradient all-reduce, clipping by the global norm).
Skipped on a single-GPU box; run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_policy_dp.py -m gpu`."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from conftest import ROOT, load_golden

pytestmark = pytest.mark.gpu


def _make_agent(g, device):
    from marllb_b200.policy import QMIXAgent
    agent = QMIXAgent(num_agents=3, state_dim=int(g["batch_states"].shape[-1]), obs_dim=int(g["batch_observations"].shape[-1]),
                      action_dim=5, hidden_dim=32, gru_dim=16, mixing_embed_dim=8, hypernet_embed_dim=16,
                      batch_size=4, max_seq_len=int(g["batch_observations"].shape[1]), device=device)
    return agent


def _full_batch(g):
    b = {k[len("batch_"):]: np.array(v) for k, v in g.items() if k.startswith("batch_")}
    T = b["observations"].shape[1]
    b["seq_lengths"] = np.full_like(b["seq_lengths"], T)     # equal valid counts on both halves
    b["dones"] = np.zeros_like(b["dones"])
    return b


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    g = load_golden("policy_qmix")
    torch.manual_seed(7)
    agent = _make_agent(g, torch.device("cuda", rank))
    b = _full_batch(g)
    B = b["observations"].shape[0]
    half = B // world
    mine = {k: v[rank * half:(rank + 1) * half] for k, v in b.items()}
    for _ in range(2):
        agent.update(batch=mine)
    flat = agent._bucket.flat_p.detach().cpu().numpy()
    q.put((rank, flat))
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_qmix_update_two_ranks_equals_global_batch():
    g = load_golden("policy_qmix")
    assert g["batch_observations"].shape[0] % 2 == 0
    torch.manual_seed(7)
    ref = _make_agent(g, torch.device("cuda", 0))
    b = _full_batch(g)
    for _ in range(2):
        ref.update(batch=b)
    want = ref._bucket.flat_p.detach().cpu().numpy()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted((q.get(timeout=300) for _ in procs), key=lambda x: x[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert np.array_equal(out[0][1], out[1][1])                     # ranks stay in lock-step, bit for bit
    np.testing.assert_allclose(out[0][1], want, rtol=2e-4, atol=2e-6)


# ---- SAC (config C4's learner): four gradient buckets, all-reduces started on the side stream (Adam.reduce_async) and
# ---- captured in a CUDA graph together with the update
def _sac_setup(device, B):
    from marllb_b200.policy import SAC_GRU_Agent
    torch.manual_seed(11)
    agent = SAC_GRU_Agent(state_dim=44, action_dim=4, hidden_dim=64, gru_dim=32, batch_size=B, device=device)
    rng = np.random.RandomState(4)
    f = lambda a: torch.as_tensor(a, dtype=torch.float32)
    batch = (f(rng.randn(B, 44)), f(np.tanh(rng.randn(B, 4))), f(rng.rand(B, 1)), f(rng.randn(B, 44)),
             f((rng.rand(B, 1) < 0.2).astype(np.float32)), f(rng.randn(1, B, 32) * 0.2))
    eps = (f(rng.randn(B, 4)), f(rng.randn(B, 4)))
    return agent, batch, eps


def _sac_params(agent):
    return torch.cat([agent.policy_optimizer.bucket.flat_p, agent.q1_optimizer.bucket.flat_p,
                      agent.q2_optimizer.bucket.flat_p, agent.log_alpha.reshape(-1)]).detach().cpu().numpy()


def _sac_worker(rank, world, port, q, graphed):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    B = 16
    agent, batch, eps = _sac_setup(dev, B)
    half = B // world
    sl = slice(rank * half, (rank + 1) * half)
    mine = tuple((t[:, sl] if t.dim() == 3 else t[sl]).contiguous().to(dev) for t in batch)
    e_next, e_new = eps[0][sl].contiguous().to(dev), eps[1][sl].contiguous().to(dev)
    upd = lambda: agent.update_parameters(1, batch=mine, eps_next=e_next, eps_new=e_new, sync_stats=False)
    if graphed:
        upd(); upd()                                           # eager (creates the NCCL communicator)
        side = torch.cuda.Stream(device=dev)
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            upd()                                              # captured, not executed: all-reduces included
        graph.replay(); graph.replay()
        torch.cuda.synchronize(dev)
        graph.reset()                      # a graph holding NCCL kernels must be gone before the communicator is
        del graph
    else:
        for _ in range(4):
            upd()
    torch.cuda.synchronize(dev)
    q.put((rank, _sac_params(agent)))
    import threading
    threading.Timer(30.0, lambda: os._exit(0)).start()      # teardown must not outlive the test
    dist.barrier()
    dist.destroy_process_group()
    os._exit(0)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("graphed", [False, True])
def test_sac_update_two_ranks_equals_global_batch(graphed):
    """Four SAC updates with the batch split over two ranks (averaging all-reduce of each bucket on the side stream;
    graphed=True: two of them replayed from a CUDA graph that holds the NCCL calls) = four updates with the whole
    batch on one rank; the two ranks end bit-identical."""
    B = 16
    ref, batch, eps = _sac_setup(torch.device("cuda", 0), B)
    dev = torch.device("cuda", 0)
    full = tuple(t.to(dev) for t in batch)
    for _ in range(4):
        ref.update_parameters(1, batch=full, eps_next=eps[0].to(dev), eps_new=eps[1].to(dev), sync_stats=False)
    want = _sac_params(ref)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 35500 + (os.getpid() % 2000) + (1 if graphed else 0)
    procs = [ctx.Process(target=_sac_worker, args=(r, 2, port, q, graphed)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted((q.get(timeout=300) for _ in procs), key=lambda x: x[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert np.array_equal(out[0][1], out[1][1])                     # replicas in lock-step, bit for bit
    np.testing.assert_allclose(out[0][1], want, rtol=3e-4, atol=3e-6)
