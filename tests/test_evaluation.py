"""the evaluation metrics of kitti_evaluation.cpp:44-146 (ground-segmentation confusion counts, over- /
under-segmentation entropy) on the device against the reference's own code (oracle/_ref/libcc_eval_ref.so: excerpts of
kitti_evaluation.cpp compiled unmodified) and against a numpy restatement that also runs where that build is absent.
Counts are exact; the entropies are sums of doubles in a different order: relative tolerance 1e-9."""
import ctypes as C
import os

import numpy as np
import pytest

from continuous_clustering_b200 import KittiEvaluation

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EVAL_REF = os.path.join(REPO, "oracle", "_ref", "libcc_eval_ref.so")
GROUND = [60, 40, 44, 48, 49, 72]  # lane-marking, road, parking, sidewalk, other-ground, terrain (kitti_loader.cpp:566-603)
OTHER = [0, 1, 10, 11, 30, 50, 51, 70, 71, 80, 81, 99, 252, 259]


def restatement(sem, ground, gt, det):
    """numpy restatement of kitti_evaluation.cpp:44-146 (pure counting; checked against the reference build below)."""
    lab = sem != 0
    g = np.isin(sem, GROUND)
    s = ground != 0
    out = {"tp": float((lab & g & s).sum()), "fn": float((lab & g & ~s).sum()), "fp": float((lab & ~g & s).sum()),
           "tn": float((lab & ~g & ~s).sum())}
    ose = 0.0
    for k in np.unique(gt[gt != 0]):  # cpp:101-118
        d = det[gt == k]
        _, c = np.unique(d, return_counts=True)
        f = c / d.size
        ose -= float((f * np.log(f)).sum())
    use = 0.0
    for k in np.unique(det[det != 0]):  # cpp:121-144
        t = gt[det == k]
        v, c = np.unique(t, return_counts=True)
        if len(v) == 1 and v[0] == 0:
            continue
        f = c / t.size
        use -= float((f * np.log(f)).sum())
    out["over_segmentation_entropy"], out["under_segmentation_entropy"] = ose, use
    return out


def reference(sem, ground, gt, det):
    lib = C.CDLL(EVAL_REF)
    lib.ev_evaluate.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    out = np.zeros(6)
    lib.ev_evaluate(sem.size, sem.ctypes.data, ground.ctypes.data, gt.ctypes.data, det.ctypes.data, out.ctypes.data)
    return dict(zip(["tp", "fn", "fp", "tn", "over_segmentation_entropy", "under_segmentation_entropy"], out.tolist()))


def frame(n, seed, n_gt=60, n_det=90, p_none=0.5):
    rng = np.random.RandomState(seed)
    sem = rng.choice(np.array(GROUND + OTHER, dtype=np.uint16), size=n).astype(np.uint16)
    ground = (np.isin(sem, GROUND) ^ (rng.uniform(size=n) < 0.1)).astype(np.uint8)
    # spatially coherent labels: detections mostly follow the ground-truth clusters, with splits and merges
    gt = rng.randint(1, n_gt + 1, size=n).astype(np.uint32)
    det = ((gt * 7 + (rng.uniform(size=n) < 0.3) * rng.randint(0, 5, size=n)) % n_det + 1).astype(np.uint32)
    gt[rng.uniform(size=n) < p_none] = 0
    det[rng.uniform(size=n) < p_none] = 0
    det[det == 5] = 0xfffffff0  # large ids (Point::id is a running counter cast to 32 bits)
    return sem, ground, gt, det


def check(a, b, what):
    for k in ("tp", "fn", "fp", "tn"):
        assert a[k] == b[k], f"{what}: {k} {a[k]} vs {b[k]}"
    for k in ("over_segmentation_entropy", "under_segmentation_entropy"):
        assert abs(a[k] - b[k]) <= 1e-9 * max(1.0, abs(a[k])), f"{what}: {k} {a[k]!r} vs {b[k]!r}"


CASES = [(0, 1), (1, 2), (5000, 3), (120000, 4)]


@pytest.mark.parametrize("n,seed", CASES)
def test_restatement_matches_reference_build(n, seed):
    if not os.path.exists(EVAL_REF):
        pytest.skip("oracle/_ref/libcc_eval_ref.so not built (needs /root/reference at build time)")
    args = frame(n, seed)
    check(reference(*args), restatement(*args), "restatement vs reference")


@pytest.mark.parametrize("n,seed", CASES)
def test_emulated_kernels_match_oracle(emu_library, n, seed):
    args = frame(n, seed)
    ev = KittiEvaluation(max_points_per_frame=1 << 17, _library=emu_library)
    got = ev.evaluate(*args)
    check(restatement(*args), got, "emulated kernels vs restatement")
    if os.path.exists(EVAL_REF):
        check(reference(*args), got, "emulated kernels vs reference")
    ev.close()


def test_degenerate_frames(emu_library):
    ev = KittiEvaluation(max_points_per_frame=4096, _library=emu_library)
    n = 1000
    z16, z8, z32 = np.zeros(n, np.uint16), np.zeros(n, np.uint8), np.zeros(n, np.uint32)
    assert ev.evaluate(z16, z8, z32, z32) == dict(tp=0.0, fn=0.0, fp=0.0, tn=0.0, over_segmentation_entropy=0.0,
                                                  under_segmentation_entropy=0.0)
    one = np.ones(n, np.uint32)
    got = ev.evaluate(np.full(n, 40, np.uint16), np.ones(n, np.uint8), one, one)  # one cluster, detected as one
    assert got["tp"] == n and got["over_segmentation_entropy"] == 0.0 and got["under_segmentation_entropy"] == 0.0
    got = ev.evaluate(z16, z8, z32, one)  # a detection without any ground-truth point is ignored (cpp:129-131)
    assert got["under_segmentation_entropy"] == 0.0
    ev.close()


@pytest.mark.gpu
@pytest.mark.parametrize("n,seed", CASES + [(240000, 9)])
def test_cuda_matches_oracle(cuda_library, n, seed):
    args = frame(n, seed, n_gt=300, n_det=500)
    ev = KittiEvaluation(max_points_per_frame=1 << 18)
    got = ev.evaluate(*args)
    check(restatement(*args), got, "cuda vs restatement")
    if os.path.exists(EVAL_REF):
        check(reference(*args), got, "cuda vs reference")
    ev.close()
