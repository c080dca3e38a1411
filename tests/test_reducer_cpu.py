"""CPU, world_size 2 over gloo: host logic of the bucketed gradient reducer (bucket membership, ready counting,
flush of never-produced gradients, mean over ranks). The device path (NCCL) is exercised by tests/test_dp_gpu.py."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank: int, world: int, port: int, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from diffulab_b200.training import GradReducer

        torch.manual_seed(0)
        params = [torch.nn.Parameter(torch.zeros(n)) for n in (5, 7, 3, 11)]
        red = GradReducer(params=params)
        assert len(red.buckets) == 4 and red.world == world
        for step in range(2):
            for p in params:
                p.grad.zero_()
            red.begin()
            # gradients arrive last-to-first; parameter 2 never receives one (reference quirk)
            for i in (3, 1, 0):
                params[i].grad += float(rank + 1) * (i + 1) + step
                red._ready(params[i])
            assert red.launched[red.bucket_of[id(params[3])]]
            assert not red.launched[red.bucket_of[id(params[2])]]
            red.finish()
            assert all(red.launched)
            for i in (3, 1, 0):
                expect = sum((r + 1) * (i + 1) + step for r in range(world)) / world
                assert torch.allclose(params[i].grad, torch.full_like(params[i].grad, expect)), (i, params[i].grad)
            assert params[2].grad.abs().max().item() == 0.0
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_grad_reducer_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_bucket_layout_reverse_order():
    """Flat-store bucketing logic on fake stores (no CUDA): contiguous slices, reverse parameter order, size cap."""
    from diffulab_b200.training import _ALIGN, GradReducer

    class FakeStore:
        def __init__(self, sizes):
            self.params = [torch.nn.Parameter(torch.zeros(n)) for n in sizes]
            self.offsets, off = [], 0
            for p in self.params:
                self.offsets.append(off)
                off += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
            self.flat_g = torch.zeros(off)

    st = FakeStore([100, 64, 300, 10, 129])
    red = GradReducer(stores=[st], bucket_mb=4 * 320 / (1024 * 1024))  # cap = 320 elements
    sizes = [b.numel() for b in red.buckets]
    assert sum(sizes) == st.flat_g.numel()
    # last parameters form the first bucket
    assert red.bucket_of[id(st.params[4])] == 0 and red.bucket_of[id(st.params[0])] == len(red.buckets) - 1
    for p, o in zip(st.params, st.offsets):
        b = red.buckets[red.bucket_of[id(p)]]
        lo = (b.data_ptr() - st.flat_g.data_ptr()) // 4
        assert lo <= o and o + p.numel() <= lo + b.numel()


def test_tail_buckets_are_small():
    """tail_bucket_mb: the buckets backward produces last (lowest offsets) are capped at the small size; everything is covered once"""
    from diffulab_b200.training import _ALIGN, GradReducer

    class FakeStore:
        def __init__(self, sizes):
            self.params = [torch.nn.Parameter(torch.zeros(n)) for n in sizes]
            self.offsets, off = [], 0
            for p in self.params:
                self.offsets.append(off)
                off += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
            self.flat_g = torch.zeros(off)

    st = FakeStore([256] * 64)
    mb = lambda n: 4 * n / (1024 * 1024)  # noqa: E731
    red = GradReducer(stores=[st], bucket_mb=mb(4096), tail_bucket_mb=mb(512))
    sizes = [b.numel() for b in red.buckets]
    assert sum(sizes) == st.flat_g.numel() and len(set(red.bucket_of.values())) == len(red.buckets)
    assert sizes[0] == 4096 and sizes[-1] == 512 and sizes[-2] == 512 and max(sizes) == 4096
    same = GradReducer(stores=[st], bucket_mb=mb(4096))
    assert [b.numel() for b in same.buckets] == [4096] * 4


def test_piece_partition_of_a_bucket():
    """peer-memory reducers: a bucket of n floats is cut into world pieces of one 64-aligned length (the last ones short or empty)
    that tile it exactly — the offsets every rank derives must agree without communication"""
    from diffulab_b200.training import _ALIGN, GradReducer

    for world in (2, 3, 4, 8):
        red = GradReducer.__new__(GradReducer)
        red.world = world
        for n in (64, 128, 64 * 7, 64 * 1000, 64 * 1001, 33_554_432, 37_748_736 + 64):
            piece, lens = red._pieces(n)
            assert piece % _ALIGN == 0 and len(lens) == world and sum(lens) == n
            assert all(0 <= v <= piece for v in lens) and all(v % 4 == 0 for v in lens)
            seen_short = False
            for v in lens:  # full pieces first, then at most one short piece, then empty ones
                if seen_short:
                    assert v == 0
                elif v < piece:
                    seen_short = True


def test_ema_schedule_host_logic():
    """EMA.plan(): copy on the first update, nothing between multiples of update_every, copy up to update_after_step, then the
    warm-up decay 1 - (1 + e / inv_gamma) ** -power clamped to [min_value, beta] (the rule restated in the class docstring)"""
    from diffulab_b200.training import EMA

    ema = EMA.__new__(EMA)
    ema.beta, ema.update_after_step, ema.update_every = 0.999, 20, 10
    ema.inv_gamma, ema.power, ema.min_value = 1.0, 2.0 / 3.0, 0.0
    ema.step, ema.initted = 0, False
    kinds = []
    for _ in range(61):
        kinds.append(ema.plan())
        ema.advance()
    assert kinds[0] == ("copy", 0.0)
    assert all(k == ("skip", 0.0) for i, k in enumerate(kinds) if i % 10 != 0)
    assert kinds[10] == ("copy", 0.0) and kinds[20] == ("copy", 0.0)
    for step in (30, 40, 50, 60):
        kind, decay = kinds[step]
        e = step + 1 - 20 - 1
        assert kind == "lerp" and abs(decay - min(1.0 - (1.0 + e) ** (-2.0 / 3.0), 0.999)) < 1e-12
    assert kinds[30][1] < kinds[40][1] < kinds[50][1] < kinds[60][1] <= 0.999
    ema.beta = 0.5
    assert ema.get_current_decay(10_000) == 0.5 and ema.get_current_decay(21) == 0.0
