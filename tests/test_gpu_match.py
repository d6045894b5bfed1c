"""GPU parity tests of the Hamming matching path (bit-exact indices and distances) through the C ABI."""
import numpy as np
import pytest

from rumi_slam_b200.synth import synthetic_frame, perturbed_descriptors, stereo_pair

pytestmark = pytest.mark.gpu


def matcher(mode=None, **kw):
    """mode: None = automatic kernel choice, "popc" / "umma" = force the LOP3+POPC or the tcgen05 (TMEM accumulator)
    top-2 kernel."""
    import os
    from rumi_slam_b200 import ORBmatcher
    if mode:
        os.environ["RUMI_MATCH"] = mode
    try:
        return ORBmatcher(**kw)
    finally:
        os.environ.pop("RUMI_MATCH", None)


def real_descriptors(oracle, seeds):
    return np.concatenate([oracle.extract(synthetic_frame(s))[1] for s in seeds])


@pytest.mark.parametrize("mode", ["popc", "umma"])
def test_top2_kats(mode):
    m = matcher(mode)
    zeros, ones = np.zeros((1, 32), np.uint8), np.full((1, 32), 255, np.uint8)
    i1, d1, d2 = m.top2(zeros, ones)
    assert (i1[0], d1[0], d2[0]) == (-1, 256, 256)            # 256 is not < the initial 256 (ORBmatcher.cc:236-240)
    one = zeros.copy(); one[0, 7] = 0x10
    i1, d1, d2 = m.top2(zeros, np.concatenate([ones, one, one, zeros]))
    assert (i1[0], d1[0], d2[0]) == (3, 0, 1)
    i1, d1, d2 = m.top2(zeros, np.concatenate([one, one]))    # duplicate rows: earliest index, d2 == d1
    assert (i1[0], d1[0], d2[0]) == (0, 1, 1)
    i1, d1, d2 = m.top2(zeros, np.zeros((0, 32), np.uint8))   # no candidates
    assert (i1[0], d1[0], d2[0]) == (-1, 256, 256)


def test_top2_kat_all_ones_index(oracle):
    # the reference scan starts at bestDist=256 with strict '<': a distance of exactly 256 never becomes best
    i1, d1, d2 = oracle.hamming_top2(np.zeros((1, 32), np.uint8), np.full((1, 32), 255, np.uint8))
    assert (i1[0], d1[0], d2[0]) == (-1, 256, 256)
    gi, gd1, gd2 = matcher().top2(np.zeros((1, 32), np.uint8), np.full((1, 32), 255, np.uint8))
    assert (gi[0], gd1[0], gd2[0]) == (-1, 256, 256)


@pytest.mark.parametrize("mode", ["popc", "umma"])
@pytest.mark.parametrize("nq,nt", [(1, 1), (37, 1000), (1000, 1000), (2049, 4099), (5000, 700), (130, 127), (129, 3000)])
def test_top2_matches_oracle(oracle, nq, nt, mode):
    base = real_descriptors(oracle, (0, 1))
    Q = perturbed_descriptors(base, nq, seed=1, flip_p=0.05)
    T = perturbed_descriptors(base[::-1].copy(), nt, seed=2, flip_p=0.05)
    T[nt // 2] = T[0]                                          # force exact ties
    i1, d1, d2 = matcher(mode).top2(Q, T)
    ri, rd1, rd2 = oracle.hamming_top2(Q, T)
    assert np.array_equal(i1, ri) and np.array_equal(d1, rd1) and np.array_equal(d2, rd2)


def test_top2_large_automatic_path_and_extreme_descriptors(oracle):
    """8192 x 20000 (large enough for the automatic choice of the tensor-core kernel) with all-zero / all-one rows,
    duplicates of the best match far apart (earliest index must win) and queries whose best distance is 256."""
    rng = np.random.default_rng(9)
    base = real_descriptors(oracle, (5, 6))
    nq, nt = 8192, 20000
    Q = perturbed_descriptors(base, nq, seed=5, flip_p=0.04)
    T = perturbed_descriptors(base, nt, seed=6, flip_p=0.04)
    Q[0] = 0; Q[1] = 255; T[10] = 0; T[11] = 255; T[19999] = 0
    T[15000] = T[123]; T[4000] = T[123]
    Q[2] = T[123]
    auto, popc, umma = matcher(), matcher("popc"), matcher("umma")
    ri, rd1, rd2 = oracle.hamming_top2_mt(Q, T)
    assert auto.top2(Q[:300], T)[0].shape == (300,) and auto.last_path() == "popc"
    for m in (auto, popc, umma):
        i1, d1, d2 = m.top2(Q, T)
        assert np.array_equal(i1, ri) and np.array_equal(d1, rd1) and np.array_equal(d2, rd2)
    assert auto.last_path() == "umma"
    assert ri[2] == 123 and rd1[2] == 0 and rd2[2] == 0 and ri[0] == 10
    # every train at distance 256 from the query: no match
    for m in (popc, umma):
        i1, d1, d2 = m.top2(np.zeros((300, 32), np.uint8), np.full((2000, 32), 255, np.uint8))
        assert (i1 == -1).all() and (d1 == 256).all() and (d2 == 256).all()


@pytest.mark.parametrize("mode", ["popc", "umma"])
def test_sharded_merge_equals_single(oracle, mode):
    import torch
    base = real_descriptors(oracle, (2, 3, 4))
    Q = torch.from_numpy(perturbed_descriptors(base, 3000, seed=3)).cuda()
    T = torch.from_numpy(perturbed_descriptors(base, 8000, seed=4)).cuda()
    m = matcher(mode)
    i1, d1, d2 = m.top2_device(Q, T)
    for shards in (2, 3, 8):
        bounds = np.linspace(0, T.shape[0], shards + 1).astype(int)
        packed = torch.empty((shards, Q.shape[0]), dtype=torch.int64, device="cuda")
        for s in range(shards):
            a, b, c = m.top2_device(Q, T[bounds[s]:bounds[s + 1]].contiguous(), t_base=int(bounds[s]))
            m.pack_device(a, b, c, out=packed[s])
        gi, g1, g2 = m.merge_device(packed, shards, Q.shape[0])
        assert torch.equal(gi, i1) and torch.equal(g1, d1) and torch.equal(g2, d2)
    ri, rd1, rd2 = oracle.hamming_top2(Q.cpu().numpy(), T.cpu().numpy())
    assert np.array_equal(i1.cpu().numpy(), ri) and np.array_equal(d1.cpu().numpy().astype(np.uint16), rd1)
    assert np.array_equal(d2.cpu().numpy().astype(np.uint16), rd2)


def test_cfg5a_full_size_bitexact(oracle):
    """BASELINE config 5a at its full size -- 40 000 x 40 000 descriptors of extracted frames, the exact problem
    bench.py times on the tcgen05 kernel -- every query against the oracle's scan (ORBmatcher.cc:253-261)."""
    import torch
    base = real_descriptors(oracle, range(20, 30))
    nq = nt = 40000
    Q = perturbed_descriptors(base, nq, seed=11, flip_p=0.05)
    T = perturbed_descriptors(base[::-1].copy(), nt, seed=12, flip_p=0.05)
    T[39999] = T[17]; T[20000] = T[17]; Q[5] = T[17]             # far-apart exact ties: earliest index wins
    m = matcher()
    i1, d1, d2 = m.top2_device(torch.from_numpy(Q).cuda(), torch.from_numpy(T).cuda())
    assert m.last_path() == "umma"
    ri, rd1, rd2 = oracle.hamming_top2_mt(Q, T)
    assert np.array_equal(i1.cpu().numpy(), ri)
    assert np.array_equal(d1.cpu().numpy().astype(np.uint16), rd1) and np.array_equal(d2.cpu().numpy().astype(np.uint16), rd2)
    assert ri[5] == 17 and rd1[5] == 0 and rd2[5] == 0


def test_cfg5b_scale_sampled_queries(oracle):
    """Config 5b scale on one GPU: 262 144 queries x 1 000 000 train rows through the tcgen05 kernel (the train set
    exceeds one 2^21-row slice budget several times); 1 500 sampled queries are checked against a brute-force scan of
    all 10^6 train rows."""
    import torch
    base = torch.from_numpy(real_descriptors(oracle, range(30, 36))).cuda()
    gen = torch.Generator(device="cuda"); gen.manual_seed(3)
    def tiled(n):
        out = base.repeat(-(-n // base.shape[0]), 1)[:n].clone()
        mask = torch.zeros_like(out)
        for bit in range(8):
            mask |= (torch.rand(out.shape, device="cuda", generator=gen) < 0.1).to(torch.uint8) << bit
        return (out ^ mask).contiguous()
    Q, T = tiled(262144), tiled(1000000)
    m = matcher()
    i1, d1, d2 = m.top2_device(Q, T)
    assert m.last_path() == "umma"
    pick = np.random.default_rng(4).choice(Q.shape[0], 1500, replace=False)
    ri, rd1, rd2 = oracle.hamming_top2_mt(Q[torch.from_numpy(pick).cuda()].cpu().numpy(), T.cpu().numpy())
    assert np.array_equal(i1.cpu().numpy()[pick], ri)
    assert np.array_equal(d1.cpu().numpy().astype(np.uint16)[pick], rd1)
    assert np.array_equal(d2.cpu().numpy().astype(np.uint16)[pick], rd2)


def test_sharded_entry_point_single_rank(oracle):
    """rumi_hamming_top2_sharded with a one-rank NCCL communicator (what the driver's single-GPU box can run): scan ->
    packed candidates -> ncclAllGather -> fold on the matcher's stream == the plain scan == the oracle."""
    import torch
    base = real_descriptors(oracle, (2, 3))
    Q = torch.from_numpy(perturbed_descriptors(base, 3000, seed=3)).cuda()
    T = torch.from_numpy(perturbed_descriptors(base, 9000, seed=4)).cuda()
    m = matcher()
    m.comm_init(m.nccl_unique_id(), 0, 1)
    for _ in range(3):                                          # repeated calls reuse the gather buffers
        i1, d1, d2 = m.top2_sharded(Q, T, 0, sync=False)
    torch.cuda.synchronize()
    ri, rd1, rd2 = oracle.hamming_top2_mt(Q.cpu().numpy(), T.cpu().numpy())
    assert np.array_equal(i1.cpu().numpy(), ri) and np.array_equal(d1.cpu().numpy().astype(np.uint16), rd1)
    assert np.array_equal(d2.cpu().numpy().astype(np.uint16), rd2)


def _nccl_worker(rank, world, port, q_np, t_np, ret):
    import os
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)       # carries only the 128-byte NCCL id
    from rumi_slam_b200 import ORBmatcher
    from rumi_slam_b200.sharding import init_matcher_comm, sharded_top2, train_shard
    m = ORBmatcher(device=rank)
    init_matcher_comm(m)
    Q = torch.from_numpy(q_np).cuda(rank)
    b, e = train_shard(len(t_np), rank, world)
    T = torch.from_numpy(t_np[b:e]).cuda(rank)
    for _ in range(4):
        i1, d1, d2 = sharded_top2(m, Q, T, b)
    torch.cuda.synchronize()
    ret.put((rank, i1.cpu().numpy(), d1.cpu().numpy().astype(np.uint16), d2.cpu().numpy().astype(np.uint16)))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_nccl_equals_single_gpu(oracle):
    """World-size-N NCCL run (every visible GPU, skipped on a one-GPU box): every rank's result of the train-sharded
    match equals the unsharded scan, on a problem large enough for the tcgen05 kernel per shard."""
    import socket
    import torch
    import torch.multiprocessing as mp
    world = torch.cuda.device_count()
    if world < 2:
        pytest.skip("needs at least 2 GPUs (run with gpurun --gpus 2)")
    base = real_descriptors(oracle, (7, 8, 9))
    Q = perturbed_descriptors(base, 16384, seed=21, flip_p=0.05)
    T = perturbed_descriptors(base[::-1].copy(), 30011, seed=22, flip_p=0.05)
    T[30000] = T[3]; T[15000] = T[3]; Q[9] = T[3]
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    procs = [ctx.Process(target=_nccl_worker, args=(r, world, port, Q, T, ret)) for r in range(world)]
    for p in procs:
        p.start()
    got = [ret.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(60)
    ri, rd1, rd2 = oracle.hamming_top2_mt(Q, T)
    for rank, i1, d1, d2 in got:
        assert np.array_equal(i1, ri) and np.array_equal(d1, rd1) and np.array_equal(d2, rd2), rank
    assert ri[9] == 3


def test_stereo_best1_matches_oracle(oracle):
    from rumi_slam_b200 import ORBextractor
    left, right = stereo_pair(1)
    ex = ORBextractor(1200, 1.2, 8, 20, 7)
    _, lk, ld = ex(left)
    _, rk, rd = ex(right)
    lk, ld, rk, rd = lk.copy(), ld.copy(), rk.copy(), rd.copy()
    sf = ex.GetScaleFactors()
    fx, bf = 435.2, 47.9                                        # upstream ORB-SLAM3 EuRoC: maxD = bf / b = fx
    best, dist = matcher().stereo_best1(lk, ld, rk, rd, sf, left.shape[0], 0.0, fx)
    rbest, rdist = oracle.stereo_best1(lk, ld, rk, rd, sf, left.shape[0], 0.0, fx)
    assert np.array_equal(best, rbest) and np.array_equal(dist, rdist)
    assert (best >= 0).sum() > 100


def test_descriptor_distance(oracle):
    from rumi_slam_b200 import ORBmatcher
    rng = np.random.default_rng(0)
    for _ in range(50):
        a, b = rng.integers(0, 256, (2, 32), dtype=np.uint8)
        assert ORBmatcher.DescriptorDistance(a, b) == oracle.descriptor_distance(a, b) == \
            int(np.unpackbits(a ^ b).sum())


@pytest.mark.parametrize("w,h,nf,seed", [(752, 480, 1200, 1), (752, 480, 1200, 2), (1241, 376, 2000, 3)])
def test_compute_stereo_matches_complete(oracle, w, h, nf, seed):
    """Frame::ComputeStereoMatches end to end (cfg 3 / cfg 4 shapes): mvuRight and mvDepth bit-exact."""
    from rumi_slam_b200 import ORBextractor
    left, right = stereo_pair(seed, w, h)
    exl, exr = ORBextractor(nf, 1.2, 8, 20, 7), ORBextractor(nf, 1.2, 8, 20, 7)     # two instances, as in Frame.cc:116-119
    _, lk, ld = exl(left)
    _, rk, rd = exr(right)
    lk, ld, rk, rd = lk.copy(), ld.copy(), rk.copy(), rd.copy()
    fx, bf = 435.2, 47.9
    u, d, n = matcher().stereo_match(exl, exr, lk, ld, rk, rd, bf, bf / fx)
    ru, rdp, rn = oracle.stereo_match(left, right, lk, ld, rk, rd, bf, bf / fx)
    assert n == rn and n > 100
    assert np.array_equal(u, ru) and np.array_equal(d, rdp)


def test_keyframe_pair_association_equals_pairwise(oracle):
    """rumi_hamming_top2_pairs (one launch for all matched key-frame pairs of a submap merge) == the oracle's top-2
    pair by pair, including ragged / empty pairs, duplicates (earliest index) and the acceptance rule."""
    from rumi_slam_b200 import ORBmatcher
    from rumi_slam_b200.synth import perturbed_descriptors
    rng = np.random.default_rng(21)
    base = rng.integers(0, 256, (3000, 32), dtype=np.uint8)
    sizes = [(1000, 1000), (1, 700), (513, 0), (0, 40), (128, 129), (257, 2000), (1200, 5), (1000, 1000)]
    A, B = [], []
    for i, (na, nb) in enumerate(sizes):
        b = base[rng.choice(len(base), nb, replace=True)] if nb else np.zeros((0, 32), np.uint8)   # duplicates
        a = perturbed_descriptors(b, na, seed=i, flip_p=0.06) if nb and na else rng.integers(0, 256, (na, 32), dtype=np.uint8)
        A.append(a); B.append(b)
    m = ORBmatcher(0.7)
    got = m.top2_pairs(A, B)
    nmatch = 0
    for (i1, d1, d2), a, b in zip(got, A, B):
        ri, r1, r2 = oracle.hamming_top2(a, b)
        assert np.array_equal(i1, ri) and np.array_equal(d1, r1) and np.array_equal(d2, r2)
    for mt, a, b in zip(m.match_keyframe_pairs(A, B), A, B):
        want, _, _ = m.match_bow(a, b)
        assert np.array_equal(mt, want)
        nmatch += int((mt >= 0).sum())
    assert nmatch > 1000
    assert m.top2_pairs([], []) == []


def test_top2_umma_random_shapes(oracle):
    """The persistent tensor-core kernel cuts the flattened (query block, train tile) grid into one range per SM: ranges that
    start / end inside a query block, several segments per CTA, ragged last tiles and blocks.  40 random shapes, forced onto
    that kernel, against the oracle scan."""
    rng = np.random.default_rng(123)
    base = real_descriptors(oracle, (2, 3, 4))
    m = matcher("umma")
    shapes = [(int(rng.integers(1, 6000)), int(rng.integers(1, 9000))) for _ in range(34)]
    shapes += [(256, 128), (257, 129), (255, 127), (512, 148 * 128), (148 * 256, 128), (300, 18945)]
    for nq, nt in shapes:
        Q = perturbed_descriptors(base, nq, seed=nq, flip_p=0.05)
        T = perturbed_descriptors(base[::-1].copy(), nt, seed=nt, flip_p=0.05)
        if nt > 3:
            T[nt - 1] = T[nt // 3]                             # a tie between far-apart rows: the earlier index must win
        i1, d1, d2 = m.top2(Q, T)
        ri, rd1, rd2 = oracle.hamming_top2(Q, T)
        assert m.last_path() == "umma"
        assert np.array_equal(i1, ri) and np.array_equal(d1, rd1) and np.array_equal(d2, rd2), (nq, nt)
