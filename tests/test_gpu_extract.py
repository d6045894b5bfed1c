"""GPU parity tests of the extraction path: every stage and the full ORBextractor::operator() output are compared
BIT-EXACTLY (keypoint order included) with the oracle on the same seeded synthetic frames, through the C ABI."""
import os

import numpy as np
import pytest

from rumi_slam_b200.synth import synthetic_frame, synthetic_batch

pytestmark = pytest.mark.gpu

SHAPES = [(640, 480, 1000), (752, 480, 1200), (1241, 376, 2000)]


def make(nf, **kw):
    from rumi_slam_b200 import ORBextractor
    return ORBextractor(nf, 1.2, 8, 20, 7, **kw)


@pytest.mark.parametrize("w,h,nf", SHAPES)
@pytest.mark.parametrize("mode", ["strip", "strip1", "strip3", "strip8", "tma", "plain"])
def test_pyramid_and_blur_bitexact(oracle, w, h, nf, mode):
    # strip = the all-level shared-memory kernel (default; stripN forces N strips per frame: different halo rows),
    # tma / plain = the per-level tile kernel with / without TMA staging (fallback for scale factors > 2)
    os.environ["RUMI_PYRAMID"] = "strip" if mode.startswith("strip") else "tiles"
    os.environ["RUMI_NO_TMA"] = "1" if mode == "plain" else "0"
    if mode[5:]:
        os.environ["RUMI_PYR_STRIPS"] = mode[5:]
    try:
        ex = make(nf)
    finally:
        os.environ["RUMI_NO_TMA"] = "0"
        os.environ.pop("RUMI_PYRAMID", None)
        os.environ.pop("RUMI_PYR_STRIPS", None)
    img = synthetic_frame(11, w, h)
    ex(img)
    ref = oracle.pyramid(img)
    got = ex.mvImagePyramid
    for l in range(8):
        assert got[l].shape == ref[l].shape
        assert np.array_equal(got[l], ref[l]), "pyramid level %d differs" % l
    blur = ex.blurred_pyramid()
    for l in range(8):
        assert np.array_equal(blur[l], oracle.blur(ref[l])), "blurred level %d differs" % l


@pytest.mark.parametrize("w,h,nf", SHAPES)
def test_fast_candidates_and_octree_bitexact(oracle, w, h, nf):
    ex = make(nf)
    img = synthetic_frame(5, w, h)
    ex(img)
    ref_py = oracle.pyramid(img)
    tb = oracle.tables(nf)
    nfallback = 0
    for l in range(8):
        cand, nfb = oracle.grid_fast(ref_py[l])
        nfallback += nfb
        got = ex.debug_candidates(l)
        assert np.array_equal(got, cand.astype(np.int32)), "FAST candidates (x,y,response,order) level %d" % l
        lh, lw = ref_py[l].shape
        sel = oracle.octree(cand, 16, lw - 16, 16, lh - 16, int(tb["quota"][l]))
        gsel = ex.debug_candidates(l, selected=True)
        assert np.array_equal(gsel, cand[sel].astype(np.int32)), "octree selection level %d" % l
    assert nfallback > 0, "synthetic frame should exercise the minThFAST fallback"


@pytest.mark.parametrize("w,h,nf", SHAPES + [(640, 480, 2000), (640, 480, 5000)])
@pytest.mark.parametrize("lap", [(0, 0), (0, 1000)])
def test_extract_matches_oracle(oracle, w, h, nf, lap):
    ex = make(nf)
    for seed in (0, 1):
        img = synthetic_frame(seed, w, h)
        mono, kps, desc = ex(img, None, lap)
        rk, rd, rmono = oracle.extract(img, nfeatures=nf, lapping=lap)
        assert mono == rmono
        assert len(kps) == len(rk)
        for f in ("x", "y", "size", "response", "octave", "class_id"):
            assert np.array_equal(kps[f], rk[f]), f
        # north_star tolerance for angles is 1e-3 rad; this implementation is bit-exact
        assert np.array_equal(kps["angle"], rk["angle"])
        assert np.array_equal(desc, rd)


def test_batch_equals_single_and_device_path(oracle):
    import torch
    n, w, h = 12, 640, 480
    frames = synthetic_batch(n, w, h, seed0=40)
    single = make(1000)
    ref = [single(frames[i]) for i in range(n)]
    ex = make(1000, max_batch=5)             # 3 chunks, the last one ragged
    kps, desc, nkp, nmono = ex.extract_batch(frames)
    for i in range(n):
        m, k, d = ref[i]
        assert nkp[i] == len(k) and nmono[i] == m
        assert np.array_equal(kps[i, :nkp[i]], k) and np.array_equal(desc[i, :nkp[i]], d)
    dk, dd, dn, dm = ex.extract_batch_device(torch.from_numpy(frames).cuda())
    dk = dk.cpu().numpy().view(np.uint8).reshape(n, -1, 28)
    for i in range(n):
        m, k, d = ref[i]
        assert int(dn[i]) == len(k) and int(dm[i]) == m
        assert np.array_equal(dk[i, :len(k)].reshape(-1).view(k.dtype), k)
        assert np.array_equal(dd[i, :len(k)].cpu().numpy(), d)
    # unaligned device input (odd row stride) takes the plain-load pyramid path
    pad = torch.zeros((n, h, w + 3), dtype=torch.uint8, device="cuda")
    pad[:, :, :w] = torch.from_numpy(frames).cuda()
    uk, ud, un, um = ex.extract_batch_device(pad[:, :, :w])
    assert torch.equal(un, dn) and torch.equal(ud, dd)


def test_edge_cases(oracle):
    ex = make(1000)
    mono, kps, desc = ex(np.zeros((0, 0), np.uint8))
    assert mono == -1 and len(kps) == 0                       # ORBextractor.cc:1017
    flat = np.full((480, 640), 77, np.uint8)
    mono, kps, desc = ex(flat)
    assert mono == 0 and len(kps) == 0                        # nkeypoints == 0 path (:1035-1036)
    rng = np.random.default_rng(3)
    noise = rng.integers(0, 256, (480, 640), dtype=np.uint8)  # stress: tens of thousands of candidates per level
    mono, kps, desc = ex(noise)
    rk, rd, rmono = oracle.extract(noise)
    assert mono == rmono and np.array_equal(kps, rk) and np.array_equal(desc, rd)
    small = synthetic_frame(9, 320, 240)                      # different shape on the same handle
    ex4 = make(500)
    from rumi_slam_b200 import RumiError
    with pytest.raises(RumiError):
        ex4(small[:100, :100])                                # level 7 smaller than one FAST cell: reference divides by 0
    ex5 = __import__("rumi_slam_b200").ORBextractor(500, 1.2, 4, 20, 7)
    mono, kps, desc = ex5(small)
    rk, rd, rmono = oracle.extract(small, nfeatures=500, nlevels=4)
    assert mono == rmono and np.array_equal(kps, rk) and np.array_equal(desc, rd)
    view = np.zeros((240, 400), np.uint8)
    view[:, :320] = small
    mono2, kps2, desc2 = ex5(view[:, :320])                   # non-contiguous rows (stride 400)
    assert np.array_equal(kps2, rk) and np.array_equal(desc2, rd)


def test_cloud_frame_compute_descriptors(oracle):
    ex = make(1000)
    img = synthetic_frame(21, 640, 480)
    _, kps, _ = oracle.extract(img)[0:3]
    kps = oracle.extract(img)[0]
    level0 = kps[kps["octave"] == 0]
    n, desc = ex.CloudFrameComputeDescriptors(img, level0)
    rc, rdesc = oracle.describe(img, level0)
    assert n == rc == len(level0)
    assert np.array_equal(desc, rdesc)
    assert ex.CloudFrameComputeDescriptors(np.zeros((0, 0), np.uint8), level0)[0] == -1


@pytest.mark.parametrize("w,h,nf,levels,scale", [(1920, 1080, 3000, 8, 1.2), (1280, 720, 1500, 6, 1.3), (320, 240, 500, 5, 1.2),
                                                 (641, 479, 1000, 8, 1.2), (333, 257, 300, 4, 1.5), (800, 600, 1000, 3, 2.0),
                                                 (1024, 768, 2000, 12, 1.1)])
def test_other_shapes_and_pyramid_parameters(oracle, w, h, nf, levels, scale):
    """Shapes / pyramid parameters the benchmark does not use: full HD, odd sizes (partial words and cells), few and
    many levels, scale factors up to 2.0 (source windows of 8 pixels per 4 outputs) -- single frames (many thin strips)
    and a chunk of 9 frames (fewer, taller strips)."""
    from rumi_slam_b200 import ORBextractor
    ex = ORBextractor(nf, scale, levels, 20, 7, max_batch=9)
    frames = synthetic_batch(9, w, h, seed0=70)
    ref = [oracle.extract(frames[i], nfeatures=nf, scale=scale, nlevels=levels) for i in range(9)]
    mono, kps, desc = ex(frames[0])
    assert mono == ref[0][2] and np.array_equal(kps, ref[0][0]) and np.array_equal(desc, ref[0][1])
    bk, bd, nkp, nmono = ex.extract_batch(frames)
    for i in range(9):
        rk, rd, rm = ref[i]
        assert nkp[i] == len(rk) and nmono[i] == rm
        assert np.array_equal(bk[i, :nkp[i]], rk) and np.array_equal(bd[i, :nkp[i]], rd)


def test_benchmark_pipeline_matches_oracle(oracle):
    """What bench.py times is what is tested: a 256-frame batch in chunks of 64 (the marching / strip pyramid, every
    workspace and stream of the pipeline in use) through BOTH public batch calls -- device-resident (2 workspaces) and
    host buffers (4 workspaces, H2D + D2H inside) -- every frame compared with the oracle, keypoint order included."""
    import torch
    from rumi_slam_b200 import ORBextractor
    n, w, h = 256, 640, 480
    frames = synthetic_batch(n, w, h, seed0=3000, unique=48)
    ref = oracle.extract_mt(frames)                          # 256 oracle extractions on the host threads
    ex = ORBextractor(1000, 1.2, 8, 20, 7, max_batch=64)
    dev_in = torch.from_numpy(frames).cuda()
    for rep in range(2):                                     # second pass: workspaces are reused
        dk, dd, dn, dm = ex.extract_batch_device(dev_in, sync=False)
    torch.cuda.synchronize()
    dk = dk.cpu().numpy().view(np.uint8).reshape(n, -1, 28)
    dd, dn, dm = dd.cpu().numpy(), dn.cpu().numpy(), dm.cpu().numpy()
    hk, hd, hn, hm = ex.extract_batch(frames)
    for i in range(n):
        rk, rd, rm = ref[i]
        assert dn[i] == len(rk) == hn[i] and dm[i] == rm == hm[i], i
        assert np.array_equal(dk[i, :len(rk)].reshape(-1).view(rk.dtype), rk), "device path, frame %d" % i
        assert np.array_equal(dd[i, :len(rk)], rd), "device path descriptors, frame %d" % i
        assert np.array_equal(hk[i, :len(rk)], rk) and np.array_equal(hd[i, :len(rk)], rd), "host path, frame %d" % i
    # ragged tail + pageable (non-pinned) host memory + a second shape on the same handle
    frames2 = synthetic_batch(70, 752, 480, seed0=3100, unique=10)
    ex2 = ORBextractor(1200, 1.2, 8, 20, 7, max_batch=64)
    k2, d2, n2, m2 = ex2.extract_batch(frames2)
    ref2 = oracle.extract_mt(frames2, nfeatures=1200)
    for i in range(70):
        rk, rd, rm = ref2[i]
        assert n2[i] == len(rk) and m2[i] == rm and np.array_equal(k2[i, :len(rk)], rk) and np.array_equal(d2[i, :len(rk)], rd), i


def test_begin_end_equals_call_and_two_extractors_in_flight(oracle):
    """rumi_orb_extract_begin / _end: the same result as the one-shot call; two handles in flight from one host thread (the
    left / right image of a stereo frame); _end without _begin is an error."""
    from rumi_slam_b200 import ORBextractor
    from rumi_slam_b200.synth import synthetic_frame
    a, b = synthetic_frame(31, 752, 480), synthetic_frame(32, 752, 480)
    exl, exr = ORBextractor(1200, 1.2, 8, 20, 7), ORBextractor(1200, 1.2, 8, 20, 7)
    for _ in range(3):
        exl.begin(a); exr.begin(b)
        ml, kl, dl = exl.end()
        mr, kr, dr = exr.end()
        for (m, k, d), img in (((ml, kl, dl), a), ((mr, kr, dr), b)):
            ok, od, om = oracle.extract(img, nfeatures=1200)
            assert m == om and np.array_equal(k, ok) and np.array_equal(d, od)
    m2, k2, d2 = exl(a)
    assert m2 == ml and np.array_equal(k2, kl) and np.array_equal(d2, dl)
    exl._pending = (a, 480, 752)
    with pytest.raises(Exception):
        exl.end()


def test_random_shapes_and_parameters(oracle):
    """16 random (width, height, nfeatures, levels, scale factor) combinations -- odd widths (partial words, partial cells),
    non-standard aspect ratios (1-4 quad-tree roots), 2-9 levels -- single frame and a 5-frame chunk against the oracle."""
    from rumi_slam_b200 import ORBextractor
    rng = np.random.default_rng(2024)
    done = 0
    while done < 16:
        w, h = int(rng.integers(240, 1300)), int(rng.integers(200, 800))
        levels, scale = int(rng.integers(2, 10)), float(rng.choice([1.1, 1.2, 1.25, 1.4]))
        nf = int(rng.integers(200, 2500))
        if min(w, h) / scale ** (levels - 1) < 80 or w / h > 3.4 or h / w > 1.6:      # the reference's own limits (cells, roots)
            continue
        frames = synthetic_batch(5, w, h, seed0=9000 + done)
        ex = ORBextractor(nf, scale, levels, 20, 7, max_batch=5)
        ref = [oracle.extract(frames[i], nfeatures=nf, scale=scale, nlevels=levels) for i in range(5)]
        mono, kps, desc = ex(frames[0])
        assert mono == ref[0][2] and np.array_equal(kps, ref[0][0]) and np.array_equal(desc, ref[0][1]), (w, h, nf, levels, scale)
        bk, bd, nkp, nmono = ex.extract_batch(frames)
        for i in range(5):
            rk, rd, rm = ref[i]
            assert nkp[i] == len(rk) and nmono[i] == rm, (w, h, nf, levels, scale, i)
            assert np.array_equal(bk[i, :nkp[i]], rk) and np.array_equal(bd[i, :nkp[i]], rd), (w, h, nf, levels, scale, i)
        ex.close()
        done += 1


def test_staged_pyramid_download(oracle):
    """rumi_orb_set_pyramid_staging: the levels that follow a single-frame call into the pinned block are the same bytes the
    device-side download returns; a batch call in between invalidates the block (the levels of ITS last chunk are served)."""
    from rumi_slam_b200 import ORBextractor
    ex = ORBextractor(1000, 1.2, 8, 20, 7)
    ex._L.rumi_orb_set_pyramid_staging(ex._h, 1)
    for seed, (w, h) in ((21, (640, 480)), (22, (641, 479)), (23, (640, 480))):
        img = synthetic_frame(seed, w, h)
        _, kps, desc = ex(img)
        ref = oracle.pyramid(img)
        got = ex.mvImagePyramid
        for l in range(8):
            assert got[l].shape == ref[l].shape and np.array_equal(got[l], ref[l]), (seed, l)
        rk, rd, _ = oracle.extract(img)
        assert np.array_equal(kps, rk) and np.array_equal(desc, rd)
    frames = synthetic_batch(3, 640, 480, seed0=500)
    ex.extract_batch(frames)
    got = ex.mvImagePyramid                                   # served by the device-side path again: the last chunk (= frame 2)
    ref = oracle.pyramid(frames[2])
    for l in range(8):
        assert np.array_equal(got[l], ref[l]), l


def test_dense_frames_take_the_large_key_paths(oracle):
    """Textured (noisy) frames put thousands of FAST candidates on a level: more than the 2048-key shared-memory buffer of
    the quad-tree's first pass.  Single frames run every level with the large buffer (loop-form bucket sort), batches defer
    dense levels to a second pass once the handle has met one, and beyond 16384 keys the global-memory network takes over
    -- all bit-exact."""
    w, h = 640, 480
    rng = np.random.default_rng(5)
    base = synthetic_frame(3, w, h).astype(np.int32)
    frames = np.stack([np.clip(base + rng.integers(-a, a + 1, base.shape), 0, 255).astype(np.uint8)
                       for a in (8, 14, 22, 30, 60, 0, 12, 40)])
    refs = [oracle.extract(f, nfeatures=1000) for f in frames]
    single = make(1000)
    dense = 0
    for f, (rk, rd, rmono) in zip(frames, refs):
        mono, kps, desc = single(f)
        dense += len(single.debug_candidates(0)) > 2048
        assert mono == rmono and np.array_equal(desc, rd)
        for fld in ("x", "y", "size", "angle", "response", "octave"):
            assert np.array_equal(kps[fld], rk[fld]), fld
    assert dense >= 6                                    # the inputs do exercise the dense paths
    ex = make(1000, max_batch=8)                         # one chunk of 8: sparse and dense levels in the same launch
    # first call: the handle has not seen a dense level yet -> one launch, dense levels sorted in global memory (and the
    # handle learns); second call: dense levels deferred to the second pass.  Same bits both times.
    launches = []
    for _ in range(2):
        ex.launch_count(reset=True)
        kps, desc, nkp, nmono = ex.extract_batch(frames)
        launches.append(ex.launch_count())
        for i, (rk, rd, rmono) in enumerate(refs):
            assert nkp[i] == len(rk) and nmono[i] == rmono
            assert np.array_equal(desc[i, :nkp[i]], rd)
            for fld in ("x", "y", "angle", "response", "octave"):
                assert np.array_equal(kps[i, :nkp[i]][fld], rk[fld]), fld
    if not (os.environ.get("RUMI_OCTREE_TWO_PASS") or os.environ.get("RUMI_OCTREE_ONE_PASS")):
        assert launches[1] == launches[0] + 1            # the second quad-tree pass
