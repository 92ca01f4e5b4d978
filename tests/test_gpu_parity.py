"""GPU parity tests proper: libsvi_gpu (through the C-ABI / ctypes) against the CPU oracle on the
same seeded inputs.  Integer / byte / index outputs must be bit-exact; xyz within 1e-5 relative
(north_star), and in practice bit-equal because the fp64 path has no fused operations."""
import numpy as np
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("built_library", "built_oracle")]

from oracle import frontend_np as o
from svi_mapper_b200 import StereoFrontend, _lib
from svi_mapper_b200.frontend import SviError
from svi_mapper_b200.synth import stereo_pair


def _tri(cams, **kw):
    cl, cr = cams
    return o.Triangulator(o.Camera(cl.width, cl.height, cl.P), o.Camera(cr.width, cr.height, cr.P), o.StereoParams(**kw))


def _compare_frame(got: dict, ref: dict):
    n = len(ref["uv_l"])
    assert len(got["uv_l"]) == n
    np.testing.assert_array_equal(got["uv_l"], ref["uv_l"])
    np.testing.assert_array_equal(got["desc_l"], ref["desc_l"])
    np.testing.assert_array_equal(got["status"], ref["status"])
    np.testing.assert_array_equal(got["dist"], ref["dist"])
    np.testing.assert_array_equal(got["idx"], ref["idx"])
    ok = ref["status"] == 0
    np.testing.assert_array_equal(got["uv_r"][ok], ref["uv_r"][ok])
    np.testing.assert_array_equal(got["desc_r"][ok], ref["desc_r"][ok])
    if ok.any():
        rel = np.linalg.norm(got["xyz"][ok] - ref["xyz"][ok], axis=1) / np.linalg.norm(ref["xyz"][ok], axis=1)
        assert rel.max() <= 1e-5          # north_star tolerance
        np.testing.assert_array_equal(got["xyz"][ok], ref["xyz"][ok])   # and in fact bit-equal


@pytest.mark.parametrize("cams_name,seed", [("kitti_cams", 0), ("vi_cams", 4000), ("kitti1112_cams", 2000)])
def test_harris_response_bit_exact(request, cams_name, seed):
    cams = request.getfixturevalue(cams_name)
    L, _ = stereo_pair(cams[0].width, cams[0].height, seed)
    with StereoFrontend(*cams) as fe:
        got = fe.harris_response(L)
    ref = o.harris_response(L)
    bad = got.view(np.uint32) != ref.view(np.uint32)
    assert bad.sum() == 0, f"{bad.sum()} response pixels differ, first at {np.argwhere(bad)[:5]}"


@pytest.mark.parametrize("max_corners", [1000, 2000])
def test_detect_matches_gftt(kitti_cams, max_corners):
    W, H = kitti_cams[0].width, kitti_cams[0].height
    imgs = np.stack([stereo_pair(W, H, s)[0] for s in (0, 1, 2)])
    rng = np.random.default_rng(5)
    centres = np.stack([rng.uniform(0, W, 300), rng.uniform(0, H, 300)], 1)
    masks = np.stack([np.full((H, W), 255, np.uint8), o.mask_active_landmarks(W, H, centres), o.mask_active_landmarks(W, H, centres[:50])])
    with StereoFrontend(*kitti_cams, max_corners=max_corners) as fe:
        got = fe.detect(imgs, masks)
        got_nomask = fe.detect(imgs[:1])
    for f in range(3):
        ref = o.gftt(imgs[f], max_corners, mask=masks[f])
        np.testing.assert_array_equal(got[f].astype(np.int32), ref)
    np.testing.assert_array_equal(got_nomask[0].astype(np.int32), o.gftt(imgs[0], max_corners))


def test_describe_and_hamming(kitti_cams):
    W, H = kitti_cams[0].width, kitti_cams[0].height
    L, R = stereo_pair(W, H, 7)
    rng = np.random.default_rng(1)
    pts = np.stack([rng.integers(0, W, 500), rng.integers(0, H, 500)], 1).astype(np.float32)
    pts[:8] = [[28, 28], [27, 28], [W - 29, H - 29], [W - 28, 40], [40, H - 28], [28.4, 30.5], [27.5, 28.5], [100.5, 99.5]]
    with StereoFrontend(*kitti_cams) as fe:
        desc, kept = fe.describe(L, pts)
        keep_ref, desc_ref = o.brief32(L, pts)
        assert np.array_equal(np.nonzero(kept)[0], keep_ref)
        np.testing.assert_array_equal(desc[kept], desc_ref)
        _, desc_r = o.brief32(R, pts)
        idx, dist = fe.match_hamming(desc_ref[:64], desc_r)
        for q in range(64):
            i, d = o.match_hamming(desc_ref[q], desc_r)
            assert (idx[q], dist[q]) == (i, d)
        # constructed ties: the first minimum wins
        t = np.repeat(desc_ref[:1], 5, 0)
        idx, dist = fe.match_hamming(desc_ref[:1], t)
        assert idx[0] == 0 and dist[0] == 0
        idx, dist = fe.match_hamming(desc_ref[:3], np.zeros((0, 32), np.uint8))
        assert (idx == -1).all() and (dist == -1).all()


@pytest.mark.parametrize("cams_name,seed,kw", [
    ("kitti_cams", 0, {}),                       # C1
    ("kitti_cams", 1000, {"max_corners": 2000}), # C2-shaped frame
    ("kitti1112_cams", 2000, {}),                # C4-shaped frame
    ("vi_cams", 4000, {}),
])
def test_stereo_frame_parity(request, cams_name, seed, kw):
    cams = request.getfixturevalue(cams_name)
    L, R = stereo_pair(cams[0].width, cams[0].height, seed)
    with StereoFrontend(*cams, **kw) as fe:
        got = fe.add_new_landmarks(L, R)
    ref = o.add_new_landmarks(L, R, _tri(cams, **kw))
    assert (ref["status"] == 0).sum() > 100      # the synthetic pair really has matches
    _compare_frame(got, ref)


def test_stereo_batch_chunks_and_masks(kitti_cams):
    """Several chunks over several lanes, with per-frame masks; every frame equals the oracle."""
    W, H = kitti_cams[0].width, kitti_cams[0].height
    pairs = [stereo_pair(W, H, 1000 + i) for i in range(7)]
    Ls, Rs = np.stack([p[0] for p in pairs]), np.stack([p[1] for p in pairs])
    rng = np.random.default_rng(3)
    masks = np.stack([o.mask_active_landmarks(W, H, np.stack([rng.uniform(0, W, 100 * i), rng.uniform(0, H, 100 * i)], 1)) for i in range(7)])
    tri = _tri(kitti_cams)
    with StereoFrontend(*kitti_cams, chunk_frames=2) as fe:
        res = fe.stereo_frames(Ls, Rs, masks)
        res2 = fe.stereo_frames(Ls, Rs, masks)    # idempotent: scratch reuse leaves no residue
    for f in range(7):
        ref = o.add_new_landmarks(Ls[f], Rs[f], tri, mask=masks[f])
        _compare_frame(res.frame(f), ref)
        _compare_frame(res2.frame(f), ref)
        assert res.n_detected[f] == len(o.gftt(Ls[f], 1000, mask=masks[f]))


def test_triangulate_right_left_queries(kitti_cams):
    W, H = kitti_cams[0].width, kitti_cams[0].height
    L, R = stereo_pair(W, H, 11)
    tri = _tri(kitti_cams)
    rng = np.random.default_rng(2)
    n = 300
    xl = rng.integers(28, W - 28, n).astype(np.float32)
    yl = rng.integers(28, H - 28, n).astype(np.float32)
    _, dl = o.brief32(L, np.stack([xl, yl], 1))
    _, dr = o.brief32(R, np.stack([xl, yl], 1))
    rng_f = rng.uniform(0.5, 90.0, n).astype(np.float32)          # fractional search ranges (tracking)
    u_tl = np.maximum(np.float32(0), (xl - np.float32(28)) - rng_f).astype(np.float32)
    v_tl = (yl - np.float32(28)).astype(np.float32)
    u_tl[:5] = xl[:5]                                               # insufficient search range
    with StereoFrontend(*kitti_cams) as fe:
        got = fe.triangulate_right(R, np.stack([u_tl, v_tl], 1), np.stack([xl, yl], 1), dl)
        for i in range(n):
            r = tri.triangulate_right(R, u_tl[i], v_tl[i], 7.0, (xl[i], yl[i]), dl[i])
            assert got["status"][i] == r["status"], i
            assert got["dist"][i] == r.get("dist", -1) and got["idx"][i] == r.get("idx", -1), i
            if r["status"] == 0:
                assert tuple(got["uv"][i]) == tuple(np.float32(v) for v in r["uv"])
                np.testing.assert_array_equal(got["xyz"][i], r["xyz"])
                np.testing.assert_array_equal(got["desc"][i], r["desc"])
        # search in LEFT for right-image points
        tl = np.stack([xl - np.float32(28), v_tl], 1).astype(np.float32)
        sr = rng_f.copy()
        sr[:3] = [0.0, -1.0, 1e-3]
        got = fe.triangulate_left(L, sr, tl, np.stack([xl, yl], 1), dr)
        for i in range(n):
            r = tri.triangulate_left(L, sr[i], tl[i, 0], tl[i, 1], 7.0, (xl[i], yl[i]), dr[i])
            assert got["status"][i] == r["status"], i
            assert got["dist"][i] == r.get("dist", -1) and got["idx"][i] == r.get("idx", -1), i
            if r["status"] == 0:
                assert tuple(got["uv"][i]) == tuple(np.float32(v) for v in r["uv"])
                np.testing.assert_array_equal(got["xyz"][i], r["xyz"])
                np.testing.assert_array_equal(got["desc"][i], r["desc"])


def test_point_in_left_closed_form(kitti_cams):
    """Known-answer check from src/runnable/triangulation_sampling.cpp:99-120: uR = uL + Du_R/Z."""
    tri = _tri(kitti_cams)
    z = np.array([0.5, 1.0, 5.0, 20.0, 80.0])
    ul = np.full(5, 700.0, np.float32)
    ur = (ul + np.float32(1) * (tri.du_r / z)).astype(np.float32)
    uvl = np.stack([ul, np.full(5, 150.0, np.float32)], 1)
    uvr = np.stack([ur, np.full(5, 150.0, np.float32)], 1)
    with StereoFrontend(*kitti_cams) as fe:
        xyz, st = fe.point_in_left(uvl, uvr)
        xyz0, st0 = fe.point_in_left([[100.0, 50.0]], [[100.0, 50.0]])
    assert (st == 0).all() and st0[0] == _lib.SVI_TRI_ZERO_DISP
    np.testing.assert_allclose(xyz[:, 2], z, rtol=1e-4)
    for i in range(5):
        s, ref = tri.point_in_left(uvl[i], uvr[i])
        np.testing.assert_array_equal(xyz[i], ref)
        # compiled-out reference assert (CTriangulator.cpp:351): X == (Z*xR - Z*cx - DuR)/f.  Its 1e-10
        # bound assumes exact disparities; with fp32 pixel coordinates it holds to ~1e-7 relative.
        alt = tri.f_inv * (xyz[i, 2] * float(ur[i]) - xyz[i, 2] * tri.pu - tri.du_r)
        assert abs(xyz[i, 0] - alt) <= 1e-6 * max(1.0, abs(alt))


def _landmarks_from_frame(ref0, ok):
    disp = (ref0["uv_l"][ok, 0] - ref0["uv_r"][ok, 0]).astype(np.float32)
    lms = [dict(xyz_w=ref0["xyz"][i], last_desc_l=ref0["desc_l"][i], last_desc_r=ref0["desc_r"][i],
                last_disparity=disp[k], size=7.0) for k, i in enumerate(ok)]
    return disp, lms


def _compare_tracks(got, ref):
    for i, r in enumerate(ref):
        assert got["stage"][i] == r["stage"], (i, got["stage"][i], r)
        assert got["status"][i] == r["status"], (i, got["status"][i], r)
        if r["stage"]:
            assert tuple(got["uv_l"][i]) == tuple(np.float32(v) for v in r["uv_l"]), i
            assert tuple(got["uv_r"][i]) == tuple(np.float32(v) for v in r["uv_r"]), i
            np.testing.assert_array_equal(got["xyz"][i], r["xyz"])
            np.testing.assert_array_equal(got["desc_l"][i], r["desc_l"])
            np.testing.assert_array_equal(got["desc_r"][i], r["desc_r"])


def test_track_manual_stage1(vi_cams):
    """Projection-exact tracking (stage 1 LEFT/RIGHT): landmarks of frame 0 re-found under a small camera translation."""
    W, H = vi_cams[0].width, vi_cams[0].height
    L, R = stereo_pair(W, H, 4000)
    tri = _tri(vi_cams)
    ref0 = o.add_new_landmarks(L, R, tri)
    ok = np.nonzero(ref0["status"] == 0)[0][:300]
    disp, lms = _landmarks_from_frame(ref0, ok)
    T = np.eye(4)
    T[0, 3] = 0.01
    for scaling in (1.0, 2.5):
        ref = o.track_manual(L, R, tri, T, lms, scaling)
        with StereoFrontend(*vi_cams) as fe:
            got = fe.track_landmarks(L, R, T, ref0["xyz"][ok], ref0["desc_l"][ok], ref0["desc_r"][ok], disp, 7.0, scaling)
        stages = np.bincount([r["stage"] for r in ref], minlength=5)
        assert stages[1] > 50
        _compare_tracks(got, ref)


def test_track_manual_stage2_window_search(vi_cams):
    """Stage 2 (GFTT inside the projection window): the next frame is the same scene moved by (3, 2) px, so the
    exact projections miss and the window search has to re-find the corners; then the LEFT image is damaged
    in a band so that stage 2 LEFT fails there and stage 2 RIGHT takes over."""
    W, H = vi_cams[0].width, vi_cams[0].height
    L, R = stereo_pair(W, H, 4000)
    tri = _tri(vi_cams)
    ref0 = o.add_new_landmarks(L, R, tri)
    ok = np.nonzero(ref0["status"] == 0)[0][:160]
    disp, lms = _landmarks_from_frame(ref0, ok)
    L1, R1 = np.roll(L, (2, 3), axis=(0, 1)), np.roll(R, (2, 3), axis=(0, 1))
    L2 = L1.copy()
    rng = np.random.default_rng(0)
    L2[:, 150:420] = rng.integers(0, 256, size=(H, 270), dtype=np.uint8)      # LEFT unusable in this band
    for (a, b), want in (((L1, R1), 3), ((L2, R1), 4)):
        ref = o.track_manual(a, b, tri, np.eye(4), lms, 1.0)
        stages = np.bincount([r["stage"] for r in ref], minlength=5)
        assert stages[want] >= 5, stages
        with StereoFrontend(*vi_cams) as fe:
            got = fe.track_landmarks(a, b, np.eye(4), ref0["xyz"][ok], ref0["desc_l"][ok], ref0["desc_r"][ok], disp, 7.0, 1.0)
        _compare_tracks(got, ref)


def test_track_manual_stage3_epipolar(vi_cams):
    """Stage 3 (epipolar-line search): stage 2 is disabled through its cut-off so the cascade reaches the line search;
    four camera motions give horizontal, slanted and radial epipolar lines (sampling along u and along v)."""
    W, H = vi_cams[0].width, vi_cams[0].height
    L, R = stereo_pair(W, H, 4000)
    tri = _tri(vi_cams, cutoff_stage2=0.0)
    ref0 = o.add_new_landmarks(L, R, tri)
    ok = np.nonzero(ref0["status"] == 0)[0][:120]
    disp, lms = _landmarks_from_frame(ref0, ok)
    for k, lm in zip(ok, lms):
        lm.update(uv_ref=ref0["uv_l"][k].astype(np.float64), ref_desc_l=ref0["desc_l"][k], T_det_l2w=np.eye(4))
    seen_u = seen_v = 0
    with StereoFrontend(*vi_cams, cutoff_stage2=0.0) as fe:
        for t in ((0.02, 0.0, 0.0), (0.015, 0.01, 0.0), (0.0, 0.0, 0.05), (0.01, 0.03, 0.02)):
            T = np.eye(4)
            T[:3, 3] = t
            ref = o.track_manual_full(L, R, tri, T, lms, 1.5)
            for lm in lms:
                pl = o.epipolar_plan(tri, T, lm["T_det_l2w"], lm["uv_ref"], lm["xyz_w"], 1.5)
                if pl["status"] == 0:
                    seen_u += pl["along_u"]
                    seen_v += not pl["along_u"]
            got = fe.track_landmarks(L, R, T, ref0["xyz"][ok], ref0["desc_l"][ok], ref0["desc_r"][ok], disp, 7.0, 1.5,
                                     uv_reference_left=ref0["uv_l"][ok], desc_reference_left=ref0["desc_l"][ok],
                                     T_left_to_world_at_detection=np.eye(4))
            stages = np.bincount([r["stage"] for r in ref], minlength=6)
            assert stages[5] > 60, stages
            _compare_tracks(got, ref)
    assert seen_u > 100 and seen_v > 100
    # no motion since detection (and stages 1-2 disabled): the essential matrix is undefined, stage 3 is skipped
    tri0 = _tri(vi_cams, cutoff_stage1=0.0, cutoff_stage2=0.0)
    ref = o.track_manual_full(L, R, tri0, np.eye(4), lms, 1.5)
    assert all(r["stage"] == 0 for r in ref) and sum(r["status"] == o.ST_EPI_NO_TRANSLATION for r in ref) > 100
    with StereoFrontend(*vi_cams, cutoff_stage1=0.0, cutoff_stage2=0.0) as fe:
        got = fe.track_landmarks(L, R, np.eye(4), ref0["xyz"][ok], ref0["desc_l"][ok], ref0["desc_r"][ok], disp, 7.0, 1.5,
                                 uv_reference_left=ref0["uv_l"][ok], desc_reference_left=ref0["desc_l"][ok],
                                 T_left_to_world_at_detection=np.eye(4))
    _compare_tracks(got, ref)


def test_device_resident_entry(kitti_cams):
    """svi_stereo_frames_device on torch-owned device memory == the host-buffer entry point."""
    import torch
    W, H = kitti_cams[0].width, kitti_cams[0].height
    pairs = [stereo_pair(W, H, 1000 + i) for i in range(5)]
    Ls, Rs = np.stack([p[0] for p in pairs]), np.stack([p[1] for p in pairs])
    n, cap = 5, 1000
    dev = torch.device("cuda:0")
    dL, dR = torch.from_numpy(Ls).to(dev), torch.from_numpy(Rs).to(dev)
    t = dict(n_kp=torch.zeros(n, dtype=torch.int32, device=dev), n_det=torch.zeros(n, dtype=torch.int32, device=dev),
             uv_l=torch.zeros(n, cap, 2, device=dev), uv_r=torch.zeros(n, cap, 2, device=dev),
             xyz=torch.zeros(n, cap, 3, dtype=torch.float64, device=dev),
             dl=torch.zeros(n, cap, 32, dtype=torch.uint8, device=dev), dr=torch.zeros(n, cap, 32, dtype=torch.uint8, device=dev),
             dist=torch.zeros(n, cap, dtype=torch.int32, device=dev), idx=torch.zeros(n, cap, dtype=torch.int32, device=dev),
             st=torch.zeros(n, cap, dtype=torch.uint8, device=dev))
    res = _lib.StereoResult(cap, t["n_kp"].data_ptr(), t["n_det"].data_ptr(), t["uv_l"].data_ptr(), t["uv_r"].data_ptr(),
                            t["xyz"].data_ptr(), t["dl"].data_ptr(), t["dr"].data_ptr(), t["dist"].data_ptr(),
                            t["idx"].data_ptr(), t["st"].data_ptr())
    with StereoFrontend(*kitti_cams, chunk_frames=2) as fe:
        fe.stereo_frames_device(dL.data_ptr(), dR.data_ptr(), W, W * H, n, res, stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        host = fe.stereo_frames(Ls, Rs)
    np.testing.assert_array_equal(t["n_kp"].cpu().numpy(), host.n_keypoints)
    for f in range(n):
        k = host.n_keypoints[f]
        np.testing.assert_array_equal(t["uv_l"][f, :k].cpu().numpy(), host.uv_left[f, :k])
        np.testing.assert_array_equal(t["st"][f, :k].cpu().numpy(), host.status[f, :k])
        ok = host.status[f, :k] == 0
        np.testing.assert_array_equal(t["xyz"][f, :k].cpu().numpy()[ok], host.xyz_left[f, :k][ok])
        np.testing.assert_array_equal(t["dr"][f, :k].cpu().numpy()[ok], host.desc_right[f, :k][ok])


def test_golden_cv2_vectors_small_frame(kitti_cams):
    """GPU vs the committed cv2 golden vectors (tests/golden/stereo_320x240.npz): response bits, GFTT corners
    with/without mask, and the full stereo frame."""
    import pathlib
    from types import SimpleNamespace
    g = dict(np.load(pathlib.Path(__file__).resolve().parent / "golden" / "stereo_320x240.npz"))
    cl = SimpleNamespace(width=320, height=240, P=kitti_cams[0].P)
    cr = SimpleNamespace(width=320, height=240, P=kitti_cams[1].P)
    L, R = g["left"], g["right"]
    with StereoFrontend(cl, cr, max_corners=300) as fe:
        resp = fe.harris_response(L)
        np.testing.assert_array_equal(resp.view(np.uint32), g["harris"].view(np.uint32))
        np.testing.assert_array_equal(fe.detect(L)[0].astype(np.int32), g["gftt300"])
        np.testing.assert_array_equal(fe.detect(L, g["mask"])[0].astype(np.int32), g["gftt300_mask"])
        got = fe.add_new_landmarks(L, R)
    ref = {k[6:]: v for k, v in g.items() if k.startswith("frame_")}
    _compare_frame(got, ref)
    with StereoFrontend(cl, cr, max_corners=2000) as fe:      # fewer corners than maxCorners: all of them
        np.testing.assert_array_equal(fe.detect(L)[0].astype(np.int32), g["gftt_all"])


def test_properties_at_batch_scale(kitti_cams):
    """Size-independent properties on a larger batch (no oracle): the reference's compiled-out invariants
    (Types.h:115-118, CTriangulator.cpp:336-343), idempotence, and independence from batch position."""
    W, H = kitti_cams[0].width, kitti_cams[0].height
    base = [stereo_pair(W, H, 1000 + i) for i in range(4)]
    n = 40
    Ls = np.stack([base[i % 4][0] for i in range(n)])
    Rs = np.stack([base[i % 4][1] for i in range(n)])
    with StereoFrontend(*kitti_cams, max_corners=2000) as fe:
        a = fe.stereo_frames(Ls, Rs)
        b = fe.stereo_frames(Ls, Rs)
    tri = _tri(kitti_cams)
    for f in range(n):
        fa, fb, f0 = a.frame(f), b.frame(f), a.frame(f % 4)
        for k in fa:
            np.testing.assert_array_equal(fa[k], fb[k])       # idempotent
            np.testing.assert_array_equal(fa[k], f0[k])       # same pair -> same result wherever it sits in the batch
        ok = fa["status"] == 0
        assert ok.sum() > 1000
        assert (fa["uv_l"][ok, 1] == fa["uv_r"][ok, 1]).all()
        d = fa["uv_l"][ok, 0] - fa["uv_r"][ok, 0]
        assert (d >= 1).all() and (d <= 60).all() and (fa["dist"][ok] < 100).all()
        z = fa["xyz"][ok, 2]
        assert (z >= tri.depth_min).all() and (z <= tri.depth_max).all()
        np.testing.assert_allclose(z, tri.du_r_flipped / d.astype(np.float64), rtol=1e-12)
        assert ((fa["uv_l"][:, 0] >= 28) & (fa["uv_l"][:, 0] < W - 28) & (fa["uv_l"][:, 1] >= 28) & (fa["uv_l"][:, 1] < H - 28)).all()
        # min-distance property of the selected corners
        p = fa["uv_l"].astype(np.int64)
        dd = ((p[:, None, :] - p[None, :, :]) ** 2).sum(-1)
        np.fill_diagonal(dd, 1 << 30)
        assert dd.min() >= 49


def test_cpp_host_facade(vi_cams, calib_dir, tmp_path):
    """The C++ facade (svi_mapper_b200/host: CParameterBase -> CStereoCamera -> CFundamentalMatcher) driven like
    CTrackerGT drives the reference: addNewLandmarks on frame 0, trackManual on frame 1; compared with the oracle."""
    import pathlib
    import subprocess
    exe = pathlib.Path(__file__).resolve().parents[1] / "svi_mapper_b200" / "host" / "facade_demo"
    assert exe.exists(), "build it with `python -m svi_mapper_b200.build`"
    W, H = vi_cams[0].width, vi_cams[0].height
    L, R = stereo_pair(W, H, 4000)
    for name, img in (("L0", L), ("R0", R)):
        img.tofile(tmp_path / f"{name}.raw")
    out = tmp_path / "out.txt"
    r = subprocess.run([str(exe), str(calib_dir / "vi_sensor_left.txt"), str(calib_dir / "vi_sensor_right.txt"), str(W), str(H),
                        str(tmp_path / "L0.raw"), str(tmp_path / "R0.raw"), str(tmp_path / "L0.raw"), str(tmp_path / "R0.raw"), str(out),
                        str(tmp_path / "seq.cloud")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lines = out.read_text().splitlines()
    # the same facade compiled the way the reference compiles (-O3 -march=native, CMakeLists.txt:51; FMA contraction is
    # what -march=native would add): the headers pin fp-contract themselves, so every printed number is identical
    from svi_mapper_b200 import build as bld
    exe_native = bld.build_host_demo(tmp_path / "facade_demo_native", flags=("-O3", "-march=native"))
    out_n = tmp_path / "out_native.txt"
    rn = subprocess.run([str(exe_native)] + [str(a) for a in (calib_dir / "vi_sensor_left.txt", calib_dir / "vi_sensor_right.txt", W, H,
                        tmp_path / "L0.raw", tmp_path / "R0.raw", tmp_path / "L0.raw", tmp_path / "R0.raw", out_n, tmp_path / "seq_n.cloud")],
                        capture_output=True, text=True)
    assert rn.returncode == 0, rn.stderr
    assert out_n.read_text() == out.read_text()
    assert (tmp_path / "seq_n.cloud").read_bytes() == (tmp_path / "seq.cloud").read_bytes()
    tri = _tri(vi_cams)
    ref = o.add_new_landmarks(L, R, tri)
    ok = np.nonzero(ref["status"] == 0)[0]
    assert lines[0] == f"NEW {len(ok)}"
    lm = [l.split() for l in lines if l.startswith("L ")]
    assert len(lm) == len(ok)
    for row, i in zip(lm, ok):
        assert [float(v) for v in row[2:6]] == [ref["uv_l"][i, 0], ref["uv_l"][i, 1], ref["uv_r"][i, 0], ref["uv_r"][i, 1]]
        assert [float(v) for v in row[6:9]] == list(ref["xyz"][i])
    lms = [dict(xyz_w=ref["xyz"][i], last_desc_l=ref["desc_l"][i], last_desc_r=ref["desc_r"][i],
                last_disparity=np.float32(ref["uv_l"][i, 0] - ref["uv_r"][i, 0]), size=7.0) for i in ok]
    for k, lm in zip(ok, lms):
        lm.update(uv_ref=ref["uv_l"][k].astype(np.float64), ref_desc_l=ref["desc_l"][k], T_det_l2w=np.eye(4))
    trk = o.track_manual_full(L, R, tri, np.eye(4), lms, 1.0)
    n_trk = sum(1 for t in trk if t["stage"] > 0)
    n1 = sum(1 for t in trk if t["stage"] in (1, 2))
    n2 = sum(1 for t in trk if t["stage"] in (3, 4))
    n3 = sum(1 for t in trk if t["stage"] == 5)
    head = [l for l in lines if l.startswith("TRACKED")][0].split()
    assert (int(head[1]), int(head[3]), int(head[5]), int(head[7])) == (n1, n2, n3, n_trk)
    assert n_trk > 0.9 * len(ok)          # static camera: nearly everything is re-found in stage 1
    t_lines = [l.split() for l in lines if l.startswith("T ")]
    tracked = [(k, t) for k, t in enumerate(trk) if t["stage"] > 0]
    for row, (k, t) in zip(t_lines, tracked):
        assert int(row[1]) == k
        assert [float(v) for v in row[2:6]] == [float(t["uv_l"][0]), float(t["uv_l"][1]), float(t["uv_r"][0]), float(t["uv_r"][1])]
        assert float(row[6]) == t["xyz"][2]
    assert [l for l in lines if l.startswith("EXC")][0] == "EXC <CTriangulator>(getPointInLEFT) zero disparity"
    # the CTrackerGT-style sequence: frame 0 detects, later frames track (all three stages get used while the claimed
    # pose drifts away from the unchanged images) and re-detect when the trigger of CTrackerGT.cpp:305 fires
    seq = [dict(zip(l.split()[2::2], map(int, l.split()[3::2]))) for l in lines if l.startswith("SEQ")]
    assert len(seq) == 6 and seq[0]["DETECTIONS"] == 1 and seq[0]["TOTAL"] == len(ok)
    assert all(s["VISIBLE"] > 100 for s in seq) and seq[-1]["DETECTIONS"] >= 2
    assert sum(s["S1"] + s["S2"] + s["S3"] for s in seq[1:]) > 1000
    # the stereo-only tracker (CTrackerSV orchestration): frame 0 has no landmarks (pose = prior, detection), from frame 1 on
    # the pose comes from CSolverStereoPosit over the stage-1/2 measurements; the scene is static, so it stays at the origin
    sv = [l.split() for l in lines if l.startswith("SV ")]
    assert len(sv) == 12
    svd = [dict(zip(r[2:16:2], map(int, r[3:16:2]))) for r in sv]
    assert svd[0]["DETECTIONS"] == 1 and svd[0]["TOTAL"] == len(ok) and sv[0][17] == "prior"
    assert all(r[17] == "posit" for r in sv[1:]), [r[17] for r in sv]
    assert all(d["S1"] > 0.9 * len(ok) for d in svd[1:3]) and all(d["VISIBLE"] > 100 for d in svd)
    assert all(abs(float(v)) < 1e-6 for r in sv for v in r[19:22])
    assert svd[-1]["DETECTIONS"] >= 2          # the trigger of CTrackerSV.cpp:468 fires again after three frames
    # the key-frame cloud of the last sequence frame (CKeyFrame::saveCloudToFile format): visible optimal landmarks with
    # their whole LEFT descriptor history, under the pose the tracker accumulated (5 steps of 1 cm along x)
    from svi_mapper_b200 import formats
    T_l2w, cloud = formats.read_cloud(tmp_path / "seq.cloud")
    np.testing.assert_allclose(T_l2w[:3, 3], [-0.05, 0.0, 0.0], atol=1e-12)
    assert 0 < len(cloud) <= seq[-1]["VISIBLE"]
    assert all(len(p["descriptors"]) >= 1 and p["uv_l"][1] == p["uv_r"][1] and p["uv_l"][0] > p["uv_r"][0] and p["xyz_camera"][2] > 0 for p in cloud)
    # SV/SVI entry points: getPoseStereoPosit (stages 1|2 + CSolverStereoPosit), then trackEpipolar's two branches
    from svi_mapper_b200 import StereoFrontend as FE
    xyz, dl, dr = ref["xyz"][ok], ref["desc_l"][ok].copy(), ref["desc_r"][ok].copy()
    disp = (ref["uv_l"][ok, 0] - ref["uv_r"][ok, 0]).astype(np.float32)
    with FE(*vi_cams) as fe:
        def run(pose, scaling, stages, sel):
            return fe.track_landmarks(L, R, pose, xyz[sel], dl[sel], dr[sel], disp[sel], 7.0, scaling, uv_reference_left=ref["uv_l"][ok][sel],
                                      desc_reference_left=ref["desc_l"][ok][sel], T_left_to_world_at_detection=np.eye(4), stages=stages)
        every = np.arange(len(ok))
        g1 = run(np.eye(4), 1.0, 3, every)
        seen = g1["stage"] > 0
        posit = [l for l in lines if l.startswith("POSIT")][0].split()
        assert posit[0] == "POSIT" and int(posit[1]) == seen.sum() and int(posit[3]) == (g1["stage"][seen] <= 2).sum()
        assert int(posit[5]) == (g1["stage"] >= 3).sum() and int(posit[7]) == seen.sum()
        matches = [(xyz[i], g1["uv_l"][i], g1["uv_r"][i]) for i in every[seen]]
        T_ref, why = o.solve_stereo_posit(np.asarray(vi_cams[0].P).reshape(3, 4), np.asarray(vi_cams[1].P).reshape(3, 4), np.eye(4), np.zeros(3), np.eye(4), matches)
        assert why is None
        np.testing.assert_allclose(np.array([float(v) for v in posit[9:21]]).reshape(3, 4), T_ref[:3], rtol=0, atol=1e-9)

        def absorb(g, sel):   # addMeasurement: the landmark's last descriptors / disparity follow the new measurement
            hit = g["stage"] > 0
            dl[sel[hit]], dr[sel[hit]] = g["desc_l"][hit], g["desc_r"][hit]
            disp[sel[hit]] = g["uv_l"][hit, 0] - g["uv_r"][hit, 0]
            return int(hit.sum())
        absorb(g1, every)
        moved = np.eye(4)
        moved[0, 3] = 0.02
        n3 = absorb(run(moved, 1.5, 4, every), every)
        epi = [l.split() for l in lines if l.startswith("EPI")]
        assert [int(epi[0][3]), int(epi[0][5]), int(epi[0][7])] == [n3, 0, n3] and n3 > 0.5 * len(ok)
        n2 = absorb(run(np.eye(4), 1.0, 2, every), every)
        assert [int(epi[1][3]), int(epi[1][5]), int(epi[1][7])] == [0, n2, n2] and n2 > 0.5 * len(ok)


def test_stress_frame_global_select_and_long_scanlines(kitti_cams):
    """C5-shaped case: 3840x1080 frame, far more than 16384 NMS candidates (global-memory selection variant),
    3000 corners, 400 px scan-line range (several 62-candidate windows per key-point).  Checker: the C oracle."""
    from types import SimpleNamespace
    from oracle import c_oracle as co
    W, H = 3840, 1080
    L, R = stereo_pair(W, H, 3000, d_max=350)
    cl = SimpleNamespace(width=W, height=H, P=kitti_cams[0].P)
    cr = SimpleNamespace(width=W, height=H, P=kitti_cams[1].P)
    cfg = co.make_config(cl, cr, max_corners=3000, search_range=400.0)
    ref = co.frame(co.stereo_frames(cfg, L, R, n_threads=co.host_threads()), 0)
    with StereoFrontend(cl, cr, max_corners=3000, search_range_px=400.0, max_candidates=131072, chunk_frames=1) as fe:
        assert not fe.config()["select_in_smem"]
        got = fe.add_new_landmarks(L, R)
    from svi_mapper_b200 import SviError
    with StereoFrontend(cl, cr, max_corners=3000, max_candidates=16384, chunk_frames=1) as small:
        with pytest.raises(SviError, match="candidate list overflow"):   # never a silent drop
            small.add_new_landmarks(L, R)
    assert len(ref["status"]) > 2500 and (ref["status"] == 0).sum() > 2000
    d = ref["uv_l"][ref["status"] == 0, 0] - ref["uv_r"][ref["status"] == 0, 0]
    assert d.max() > 124          # matches beyond the second window really occur
    _compare_frame(got, ref)


def test_edge_cases_empty_flat_masked_padded(kitti_cams):
    """Empty and degenerate inputs: zero frames, a flat image (no corners at all), a fully masked frame, rows
    padded beyond the width (pitch > W), capacity / argument errors reported loudly."""
    from svi_mapper_b200 import SviError
    import ctypes as C
    W, H = kitti_cams[0].width, kitti_cams[0].height
    L, R = stereo_pair(W, H, 21)
    flat = np.full((H, W), 127, np.uint8)
    tri = _tri(kitti_cams)
    with StereoFrontend(*kitti_cams, chunk_frames=2) as fe:
        # flat image: Harris response is identically zero -> no candidates, no key-points; neighbours unaffected
        res = fe.stereo_frames(np.stack([L, flat, L]), np.stack([R, flat, R]))
        assert res.n_keypoints[1] == 0 and res.n_detected[1] == 0
        ref = o.add_new_landmarks(L, R, tri)
        _compare_frame(res.frame(0), ref)
        _compare_frame(res.frame(2), ref)
        assert len(o.gftt(flat, 1000)) == 0
        # fully masked frame: minMaxLoc over an empty mask, nothing detected
        zero_mask = np.zeros((H, W), np.uint8)
        res = fe.stereo_frames(np.stack([L, L]), np.stack([R, R]), np.stack([zero_mask, np.full((H, W), 255, np.uint8)]))
        assert res.n_keypoints[0] == 0
        _compare_frame(res.frame(1), ref)
        assert fe.detect(L, zero_mask)[0].shape == (0, 2)
        # zero frames is a no-op
        out = fe.stereo_frames(np.zeros((0, H, W), np.uint8), np.zeros((0, H, W), np.uint8))
        assert out.n_keypoints.shape == (0,)
        # padded rows: pitch 1280 > W, frame stride = pitch * H, through the raw C-ABI call
        pitch = 1280
        Lp, Rp = np.zeros((2, H, pitch), np.uint8), np.zeros((2, H, pitch), np.uint8)
        Lp[:, :, :W], Rp[:, :, :W] = L, R
        Lp[:, :, W:], Rp[:, :, W:] = 255, 0     # garbage in the padding must not matter
        cap = 1000
        o_ = dict(n_kp=np.zeros(2, np.int32), n_det=np.zeros(2, np.int32), uv_l=np.zeros((2, cap, 2), np.float32),
                  uv_r=np.zeros((2, cap, 2), np.float32), xyz=np.zeros((2, cap, 3)), dl=np.zeros((2, cap, 32), np.uint8),
                  dr=np.zeros((2, cap, 32), np.uint8), dist=np.zeros((2, cap), np.int32), idx=np.zeros((2, cap), np.int32),
                  st=np.zeros((2, cap), np.uint8))
        r = _lib.StereoResult(cap, *(o_[k].ctypes.data for k in ("n_kp", "n_det", "uv_l", "uv_r", "xyz", "dl", "dr", "dist", "idx", "st")))
        fe.stereo_frames_raw(Lp.ctypes.data, Rp.ctypes.data, pitch, pitch * H, 2, r)
        for f in range(2):
            k = o_["n_kp"][f]
            got = dict(uv_l=o_["uv_l"][f, :k], uv_r=o_["uv_r"][f, :k], xyz=o_["xyz"][f, :k], desc_l=o_["dl"][f, :k],
                       desc_r=o_["dr"][f, :k], dist=o_["dist"][f, :k], idx=o_["idx"][f, :k], status=o_["st"][f, :k])
            _compare_frame(got, ref)
        # loud errors: output capacity below maxCorners, pitch below the width, unknown parameter
        r_small = _lib.StereoResult(10, *(o_[k].ctypes.data for k in ("n_kp", "n_det", "uv_l", "uv_r", "xyz", "dl", "dr", "dist", "idx", "st")))
        with pytest.raises(SviError, match="capacity_per_frame"):
            fe.stereo_frames_raw(Lp.ctypes.data, Rp.ctypes.data, pitch, pitch * H, 2, r_small)
        with pytest.raises(SviError, match="bad argument"):
            fe.stereo_frames_raw(Lp.ctypes.data, Rp.ctypes.data, W - 1, pitch * H, 2, r)
        with pytest.raises(SviError, match="max_queries"):
            fe.describe(L, np.zeros((20000, 2), np.float32))
    with pytest.raises(TypeError):
        StereoFrontend(*kitti_cams, not_a_parameter=1)
    # tracking with zero landmarks and with every landmark outside the field of view
    with StereoFrontend(*kitti_cams) as fe:
        out = fe.track_landmarks(L, R, np.eye(4), np.zeros((0, 3)), np.zeros((0, 32), np.uint8), np.zeros((0, 32), np.uint8),
                                 np.zeros(0, np.float32), 7.0, 1.0)
        assert out["status"].shape == (0,)
        far = np.array([[1000.0, 0.0, 5.0], [-1000.0, 3.0, 5.0]])
        out = fe.track_landmarks(L, R, np.eye(4), far, np.zeros((2, 32), np.uint8), np.zeros((2, 32), np.uint8),
                                 np.ones(2, np.float32), 7.0, 1.0)
        assert (out["stage"] == 0).all() and (out["status"] == _lib.SVI_TRK_OUT_OF_FOV).all()


def test_harris_on_flat_and_saturated_content(kitti_cams):
    """Image content where OpenCV's fp64 RUNNING box sums lose exactness (flat / saturated regions next to strong
    edges): the GPU evaluates the same sums tile-locally.  Pinned here: (1) the GPU response equals the exact fp64
    window sum rounded once, bit for bit; (2) against OpenCV's operation order it differs only by running-sum
    residues below 1e-30 in flat regions (where the true sum is 0) or in the last bits (<= 1e-5 relative); (3) key-points, descriptors,
    matches and statuses still equal the OpenCV-order oracle."""
    import cv2
    from oracle import c_oracle as co
    W, H = kitti_cams[0].width, kitti_cams[0].height
    cfg = co.make_config(kitti_cams[0], kitti_cams[1])

    def variants(seed):
        L, R = stereo_pair(W, H, seed)
        a, b = L.copy(), R.copy()
        a[:120], b[:120], a[300:], b[300:] = 255, 255, 0, 0
        yield a, b                                                      # saturated sky / black ground
        yield (L // 32 * 32).astype(np.uint8), (R // 32 * 32).astype(np.uint8)   # posterised plateaus
        d, e = L.copy(), R.copy()
        cv2.rectangle(d, (200, 80), (500, 300), 255, -1)
        cv2.rectangle(e, (180, 80), (480, 300), 255, -1)
        cv2.line(d, (0, 0), (W - 1, H - 1), 0, 3)
        cv2.line(e, (0, 0), (W - 1, H - 1), 0, 3)
        yield d, e                                                      # high-contrast shapes
    n_diff = 0
    with StereoFrontend(*kitti_cams) as fe:
        for seed in (1, 9, 10):
            for L, R in variants(seed):
                g = fe.harris_response(L)
                exact = o.harris_response(L, box=o.box7_exact)
                np.testing.assert_array_equal(g.view(np.uint32), exact.view(np.uint32))
                cv_order = co.harris_response(L)
                bad = g.view(np.uint32) != cv_order.view(np.uint32)
                n_diff += int(bad.sum())
                if bad.any():
                    residue = (g[bad] == 0) & (np.abs(cv_order[bad]) < 1e-30)
                    # a last-bit difference of a box sum shows up as a few ulp of R = det - k*tr^2 (cancellation)
                    last_bits = np.abs(g[bad].astype(np.float64) - cv_order[bad]) <= 1e-5 * np.abs(cv_order[bad].astype(np.float64))
                    assert (residue | last_bits).all()
                ref = co.frame(co.stereo_frames(cfg, L, R), 0)
                _compare_frame(fe.add_new_landmarks(L, R), ref)
    assert n_diff > 0      # the content really triggers the effect


def test_epipolar_band_matcher(kitti_cams):
    """Optional sparse-band mode (C5 wording): key-points of both images matched inside a row band with a
    disparity window; first arg-min, distance and second-best distance equal the oracle, incl. ties and empty bands."""
    W, H = kitti_cams[0].width, kitti_cams[0].height
    L, R = stereo_pair(W, H, 31)
    kl, kr = o.gftt(L, 2000), o.gftt(R, 2000)
    keep_l, dl = o.brief32(L, kl.astype(np.float32))
    keep_r, dr = o.brief32(R, kr.astype(np.float32))
    xl, xr = kl[keep_l].astype(np.float32), kr[keep_r].astype(np.float32)
    dr[5] = dr[6] = dl[0]                                   # ties: the first admissible index wins
    xr[5] = xr[6] = xl[0] - np.float32([10, 0])
    with StereoFrontend(*kitti_cams) as fe:
        for band, dmin, dmax in ((1.0, 1.0, 60.0), (3.0, 0.0, 1e9), (0.0, 1.0, 60.0)):
            gi, gd, gs = fe.match_epipolar(dl, xl, dr, xr, band, dmin, dmax)
            ri, rd, rs = o.match_epipolar(dl, xl, dr, xr, band, dmin, dmax)
            np.testing.assert_array_equal(gi, ri)
            np.testing.assert_array_equal(gd, rd)
            np.testing.assert_array_equal(gs, rs)
        assert gi[0] == 5 and gd[0] == 0 and gs[0] == 0
        assert (ri >= 0).sum() > 200 and (ri < 0).sum() > 0
        gi, gd, gs = fe.match_epipolar(dl[:4], xl[:4], np.zeros((0, 32), np.uint8), np.zeros((0, 2), np.float32))
        assert (gi == -1).all() and (gd == -1).all() and (gs == -1).all()


def test_fast_detector_mode(kitti_cams):
    """Optional FAST-9/16 detector mode: corners equal the cv2-pinned oracle in OpenCV's output
    order, with and without non-max suppression and masks; the stereo pipeline runs on them; exceeding max_corners is
    a loud capacity error (cv::FAST has no maxCorners cut)."""
    from svi_mapper_b200 import SviError
    W, H = kitti_cams[0].width, kitti_cams[0].height
    L, R = stereo_pair(W, H, 0)
    tri = _tri(kitti_cams)
    for thr, nm in ((10, 1), (20, 0), (20, 1)):
        ref_xy, _ = o.fast9_16(L, thr, bool(nm))
        assert 200 < len(ref_xy) < 60000
        with StereoFrontend(*kitti_cams, detector=1, fast_threshold=thr, fast_nonmax=nm, max_corners=60000, max_candidates=65536,
                            chunk_frames=2) as fe:
            got = fe.detect(np.stack([L, R]))
            np.testing.assert_array_equal(got[0].astype(np.int32), ref_xy)
            np.testing.assert_array_equal(got[1].astype(np.int32), o.fast9_16(R, thr, bool(nm))[0])
            mask = o.mask_active_landmarks(W, H, ref_xy[::7].astype(np.float32))
            np.testing.assert_array_equal(fe.detect(L, mask)[0].astype(np.int32), ref_xy[mask[ref_xy[:, 1], ref_xy[:, 0]] != 0])
            if nm and thr == 20:
                res = fe.add_new_landmarks(L, R)        # the rest of the path on FAST corners
                keep, desc_l = o.brief32(L, ref_xy.astype(np.float32))
                np.testing.assert_array_equal(res["uv_l"].astype(np.int32), ref_xy[keep])
                np.testing.assert_array_equal(res["desc_l"], desc_l)
                for u in range(0, len(keep), 9):
                    x, y = np.float32(ref_xy[keep[u], 0]), np.float32(ref_xy[keep[u], 1])
                    r = tri.triangulate_right(R, max(np.float32(0), x - np.float32(60) - np.float32(28)), y - np.float32(28), 7.0, (x, y), desc_l[u])
                    assert res["status"][u] == r["status"] and res["dist"][u] == r.get("dist", -1)
    with StereoFrontend(*kitti_cams, detector=1, fast_threshold=20, max_corners=1000, max_candidates=65536) as fe:
        with pytest.raises(SviError, match="FAST found more corners"):
            fe.detect(L)


def test_track_stage_subsets(vi_cams):
    """svi_track_landmarks_stages: the pieces of the cascade the SV/SVI trackers call on their own --
    getPoseStereoPosit = stages 1|2, trackEpipolar = stage 3 alone (camera moved since detection, no field-of-view
    gate) or stage 2 alone behind the gate (CFundamentalMatcher.cpp:338-757, :760-1332)."""
    W, H = vi_cams[0].width, vi_cams[0].height
    L, R = stereo_pair(W, H, 4000)
    tri = _tri(vi_cams)
    ref0 = o.add_new_landmarks(L, R, tri)
    ok = np.nonzero(ref0["status"] == 0)[0][:90]
    disp, lms = _landmarks_from_frame(ref0, ok)
    for k, lm in zip(ok, lms):
        lm.update(uv_ref=ref0["uv_l"][k].astype(np.float64), ref_desc_l=ref0["desc_l"][k], T_det_l2w=np.eye(4))
    L1, R1 = np.roll(L, (2, 3), axis=(0, 1)), np.roll(R, (2, 3), axis=(0, 1))
    T = np.eye(4)
    T[:3, 3] = (0.015, 0.01, 0.0)
    far = np.eye(4)
    far[0, 3] = 3.0          # pushes some projections out of the field of view
    with StereoFrontend(*vi_cams) as fe:
        def run(a, b, pose, scaling, stages):
            return fe.track_landmarks(a, b, pose, ref0["xyz"][ok], ref0["desc_l"][ok], ref0["desc_r"][ok], disp, 7.0, scaling,
                                      uv_reference_left=ref0["uv_l"][ok], desc_reference_left=ref0["desc_l"][ok],
                                      T_left_to_world_at_detection=np.eye(4), stages=stages)
        seen = np.zeros(6, int)
        for a, b, pose, scaling, stages in ((L, R, T, 1.0, 1), (L1, R1, np.eye(4), 1.0, 2), (L1, R1, np.eye(4), 1.0, 3),
                                            (L, R, T, 1.5, 4), (L, R, far, 1.5, 4), (L, R, far, 1.0, 2), (L1, R1, T, 1.5, 7)):
            ref = o.track_stages(a, b, tri, pose, lms, scaling, stages)
            got = run(a, b, pose, scaling, stages)
            _compare_tracks(got, ref)
            seen += np.bincount([r["stage"] for r in ref], minlength=6)
            if stages == 2:
                assert not any(r["stage"] in (1, 2, 5) for r in ref)
            if stages == 4:
                assert not any(r["stage"] in (1, 2, 3, 4) for r in ref)
        assert seen[1] > 20 and seen[3] > 20 and seen[5] > 40, seen
        with pytest.raises(Exception):
            run(L, R, T, 1.0, 0)


@pytest.mark.parametrize("min_distance,max_corners", [(12.0, 500), (15.0, 900), (15.0, 1150), (15.0, 2000), (7.0, 3000), (3.0, 400)])
def test_detect_priority_cut_and_fallback(kitti_cams, min_distance, max_corners):
    """The selection kernel peels only the top-priority part of the candidate list (about 2.6 x maxCorners candidates)
    and falls back to the whole list when that part yields fewer than maxCorners corners.  Large minimum distances
    make the accepted fraction small, so those settings force the fallback ((15, 2000) ends below maxCorners even on the
    whole list); (7, 3000) needs every candidate anyway and (3, 400) stays on the cut list.  All must equal cv::goodFeaturesToTrack's sequential greedy result."""
    W, H = kitti_cams[0].width, kitti_cams[0].height
    L, _ = stereo_pair(W, H, 1000)
    mask = o.mask_active_landmarks(W, H, np.stack([np.linspace(40, W - 40, 60), np.linspace(30, H - 30, 60)], 1).astype(np.float32))
    with StereoFrontend(*kitti_cams, max_corners=max_corners, min_distance=min_distance) as fe:
        for m in (None, mask):
            ref = o.gftt(L, max_corners, 0.01, min_distance, m)
            got = fe.detect(L, m)[0]
            assert len(ref) > 50
            np.testing.assert_array_equal(got.astype(np.int32), np.asarray(ref, np.int32).reshape(-1, 2))


def test_small_call_path_matches_batch_path(kitti_cams):
    """Calls of one or two pairs (the tracker's per-frame use) go through a pinned bounce buffer and spread the
    key-points over more warps; they must give what the chunked batch path gives -- also with masks, with an output
    capacity above maxCorners and with padded image rows."""
    W, H = kitti_cams[0].width, kitti_cams[0].height
    pairs = [stereo_pair(W, H, 1000 + i) for i in range(5)]
    Ls, Rs = np.stack([p[0] for p in pairs]), np.stack([p[1] for p in pairs])
    rng = np.random.default_rng(9)
    masks = np.stack([o.mask_active_landmarks(W, H, np.stack([rng.uniform(0, W, 150), rng.uniform(0, H, 150)], 1)) for _ in range(5)])
    with StereoFrontend(*kitti_cams) as fe:
        big = fe.stereo_frames(Ls, Rs, masks)                       # 5 frames: batch path
        for f0, n in ((0, 1), (1, 2), (3, 2)):
            small = fe.stereo_frames(Ls[f0:f0 + n], Rs[f0:f0 + n], masks[f0:f0 + n], capacity=1300)
            for k in range(n):
                a, b = small.frame(k), big.frame(f0 + k)
                assert small.n_detected[k] == big.n_detected[f0 + k]
                ok = b["status"] == 0
                for key in a:   # the RIGHT-side fields of a key-point without a match are unspecified
                    sel = ok if key in ("uv_r", "xyz", "desc_r") else slice(None)
                    np.testing.assert_array_equal(a[key][sel], b[key][sel])
        # padded rows (pitch > W) through the one-pair call, raw C-ABI arguments
        from svi_mapper_b200 import _lib
        pitch = W + 23
        padL, padR = np.zeros((H, pitch), np.uint8), np.zeros((H, pitch), np.uint8)
        padL[:, :W], padR[:, :W] = Ls[0], Rs[0]
        cap = fe.max_corners
        o_ = dict(n_kp=np.zeros(1, np.int32), n_det=np.zeros(1, np.int32), uv_l=np.zeros((cap, 2), np.float32), uv_r=np.zeros((cap, 2), np.float32),
                  xyz=np.zeros((cap, 3)), dl=np.zeros((cap, 32), np.uint8), dr=np.zeros((cap, 32), np.uint8), dist=np.zeros(cap, np.int32),
                  idx=np.zeros(cap, np.int32), st=np.zeros(cap, np.uint8))
        r = _lib.StereoResult(cap, *(a.ctypes.data for a in (o_["n_kp"], o_["n_det"], o_["uv_l"], o_["uv_r"], o_["xyz"], o_["dl"], o_["dr"],
                                                              o_["dist"], o_["idx"], o_["st"])))
        fe.stereo_frames_raw(padL.ctypes.data, padR.ctypes.data, pitch, pitch * H, 1, r)
        ref = fe.stereo_frames(Ls[:1], Rs[:1]).frame(0)
        n = int(o_["n_kp"][0])
        assert n == len(ref["status"])
        for key, arr in (("uv_l", o_["uv_l"]), ("uv_r", o_["uv_r"]), ("xyz", o_["xyz"]), ("desc_l", o_["dl"]), ("desc_r", o_["dr"]),
                         ("dist", o_["dist"]), ("idx", o_["idx"]), ("status", o_["st"])):
            sel = (ref["status"] == 0) if key in ("uv_r", "xyz", "desc_r") else slice(None)
            np.testing.assert_array_equal(arr[:n][sel], ref[key][sel])


def test_pinned_caller_buffers_are_read_in_place(kitti_cams, vi_cams):
    """Page-locked caller images (a camera driver's DMA buffers) skip the library's pinned mirror: the one-pair call, the
    two-pair call with padded rows and the tracking call give byte-identical results from pinned and from pageable memory."""
    import torch
    W, H = kitti_cams[0].width, kitti_cams[0].height
    pairs = [stereo_pair(W, H, 1200 + i) for i in range(2)]
    Ls, Rs = np.stack([p[0] for p in pairs]), np.stack([p[1] for p in pairs])
    pitch = W + 7
    pad = torch.zeros((2, 2, H, pitch), dtype=torch.uint8).pin_memory()
    pad[0, :, :, :W], pad[1, :, :, :W] = torch.from_numpy(Ls), torch.from_numpy(Rs)
    pin = torch.stack([torch.from_numpy(Ls), torch.from_numpy(Rs)]).pin_memory()
    with StereoFrontend(*kitti_cams) as fe:
        ref = fe.stereo_frames(Ls, Rs)
        got = fe.stereo_frames(pin[0].numpy(), pin[1].numpy())
        one = fe.stereo_frames(pin[0, 1].numpy(), pin[1, 1].numpy())
        def same(a, b):   # the RIGHT-side fields of a key-point without a match are unspecified
            ok = b["status"] == 0
            for key, v in b.items():
                sel = ok if key in ("uv_r", "xyz", "desc_r") else slice(None)
                np.testing.assert_array_equal(a[key][sel], v[sel], err_msg=key)
        for k in range(2):
            same(got.frame(k), ref.frame(k))
        same(one.frame(0), ref.frame(1))
        cap = fe.max_corners
        o_ = dict(n_kp=np.zeros(2, np.int32), n_det=np.zeros(2, np.int32), uv_l=np.zeros((2, cap, 2), np.float32), uv_r=np.zeros((2, cap, 2), np.float32),
                  xyz=np.zeros((2, cap, 3)), dl=np.zeros((2, cap, 32), np.uint8), dr=np.zeros((2, cap, 32), np.uint8), dist=np.zeros((2, cap), np.int32),
                  idx=np.zeros((2, cap), np.int32), st=np.zeros((2, cap), np.uint8))
        r = _lib.StereoResult(cap, *(a.ctypes.data for a in (o_["n_kp"], o_["n_det"], o_["uv_l"], o_["uv_r"], o_["xyz"], o_["dl"], o_["dr"],
                                                              o_["dist"], o_["idx"], o_["st"])))
        fe.stereo_frames_raw(pad[0].data_ptr(), pad[1].data_ptr(), pitch, pitch * H, 2, r)
        for k in range(2):
            f = ref.frame(k)
            n = int(o_["n_kp"][k])
            assert n == len(f["status"])
            np.testing.assert_array_equal(o_["dl"][k, :n], f["desc_l"])
            np.testing.assert_array_equal(o_["st"][k, :n], f["status"])
            np.testing.assert_array_equal(o_["dist"][k, :n], f["dist"])
    # tracking call: the same landmarks against pinned and pageable frames
    Wv, Hv = vi_cams[0].width, vi_cams[0].height
    L0, R0 = stereo_pair(Wv, Hv, 4100)
    pv = torch.stack([torch.from_numpy(L0), torch.from_numpy(R0)]).pin_memory()
    with StereoFrontend(*vi_cams) as fe:
        f0 = fe.stereo_frames(L0, R0).frame(0)
        ok = np.nonzero(f0["status"] == 0)[0][:400]
        disp = (f0["uv_l"][ok, 0] - f0["uv_r"][ok, 0]).astype(np.float32)
        L1, R1 = np.roll(L0, (1, 2), axis=(0, 1)), np.roll(R0, (1, 2), axis=(0, 1))   # stage 2 has work too
        pv[0], pv[1] = torch.from_numpy(L1), torch.from_numpy(R1)
        args = (np.eye(4), f0["xyz"][ok], f0["desc_l"][ok], f0["desc_r"][ok], disp, 7.0, 1.0)
        a = fe.track_landmarks(L1, R1, *args)
        b = fe.track_landmarks(pv[0].numpy(), pv[1].numpy(), *args)
        hit = a["stage"] > 0
        assert int(hit.sum()) > 100
        np.testing.assert_array_equal(a["stage"], b["stage"])
        np.testing.assert_array_equal(a["status"], b["status"])
        for key in ("uv_l", "uv_r", "xyz", "desc_l", "desc_r"):
            np.testing.assert_array_equal(a[key][hit], b[key][hit], err_msg=key)


def test_large_batch_against_c_oracle(kitti_cams):
    """160 distinct pairs (more than two chunks on every lane pattern), maxCorners 2000: every frame of the batch path equals
    the C restatement of the reference run on all host threads -- key-points, descriptors, matches, statuses bit for bit."""
    from oracle import c_oracle as co
    from svi_mapper_b200.synth import stereo_batch_torch
    import torch
    W, H = kitti_cams[0].width, kitti_cams[0].height
    n = 160
    dL, dR = stereo_batch_torch(n, W, H, seed=7000, device=torch.device("cuda", 0))
    Ls, Rs = dL.cpu().numpy(), dR.cpu().numpy()
    cfg = co.make_config(kitti_cams[0], kitti_cams[1], max_corners=2000)
    ref = co.stereo_frames(cfg, Ls, Rs, n_threads=co.host_threads())
    with StereoFrontend(*kitti_cams, max_corners=2000) as fe:
        got = fe.stereo_frames(Ls, Rs)
    total = 0
    for f in range(n):
        r, g_ = co.frame(ref, f), got.frame(f)
        assert len(r["status"]) == len(g_["status"]) > 1000
        for k in ("uv_l", "desc_l", "status", "dist", "idx"):
            np.testing.assert_array_equal(g_[k], r[k], err_msg=f"frame {f} {k}")
        ok = r["status"] == 0
        np.testing.assert_array_equal(g_["uv_r"][ok], r["uv_r"][ok])
        np.testing.assert_array_equal(g_["desc_r"][ok], r["desc_r"][ok])
        np.testing.assert_array_equal(g_["xyz"][ok], r["xyz"][ok])
        total += int(ok.sum())
    assert total > 150 * n


@pytest.mark.parametrize("split,pre,binned", [("1", "0", "0"), ("2", "0", "0"), ("1", "1", "0"), ("2", "1", "0"), ("1", "1", "2")])
def test_matcher_variants_agree_with_c_oracle(kitti_cams, monkeypatch, split, pre, binned):
    """The scan-line matcher exists in five shapes (one or two warps per key-point; LEFT descriptors gathered inside the
    matcher or produced by describe_left_kernel ahead of it; key-points binned by position with one shared tile per bin)
    and the library picks one per launch.  Each of them, forced through the tuning knobs, equals the C restatement bit for
    bit on a batch large enough for the batch path (64 frames x 2000 slots) and on a single pair (the small-call path)."""
    from oracle import c_oracle as co
    from svi_mapper_b200.synth import stereo_batch_torch
    import torch
    monkeypatch.setenv("SVI_MATCH_SPLIT", split)
    monkeypatch.setenv("SVI_MATCH_PRE", pre)
    monkeypatch.setenv("SVI_MATCH_BINNED", binned)
    W, H = kitti_cams[0].width, kitti_cams[0].height
    n = 70   # one full chunk + a ragged one
    dL, dR = stereo_batch_torch(n, W, H, seed=7100, device=torch.device("cuda", 0))
    Ls, Rs = dL.cpu().numpy(), dR.cpu().numpy()
    cfg = co.make_config(kitti_cams[0], kitti_cams[1], max_corners=2000)
    ref = co.stereo_frames(cfg, Ls, Rs, n_threads=co.host_threads())
    with StereoFrontend(*kitti_cams, max_corners=2000) as fe:
        got = fe.stereo_frames(Ls, Rs)
        one = fe.stereo_frames(Ls[3], Rs[3])
    for f in list(range(n)) + [-1]:
        r, g_ = co.frame(ref, 3 if f < 0 else f), (one.frame(0) if f < 0 else got.frame(f))
        assert len(r["status"]) == len(g_["status"]) > 1000
        for k in ("uv_l", "desc_l", "status", "dist", "idx"):
            np.testing.assert_array_equal(g_[k], r[k], err_msg=f"frame {f} {k}")
        ok = r["status"] == 0
        for k in ("uv_r", "desc_r", "xyz"):
            np.testing.assert_array_equal(g_[k][ok], r[k][ok], err_msg=f"frame {f} {k}")


@pytest.mark.parametrize("cams_name,max_corners,search_range", [("vi_cams", 2000, 60.0), ("kitti1112_cams", 1000, 60.0),
                                                                ("kitti_cams", 2000, 37.5), ("kitti_cams", 1500, 60.75), ("vi_cams", 1000, 61.5),
                                                                ("kitti_cams", 1000, 60.0)])
def test_binned_matcher_geometries(request, monkeypatch, cams_name, max_corners, search_range):
    """The batch path sorts the key-points into image bins and matches each bin from one shared tile (binned.cuh).  Image
    sizes that are no multiple of the bin size, sparse bins (maxCorners 1000), fractional and short search ranges and the
    longest range the tile is laid out for -- and one past it (61.5: pool of 63, the library must fall back to the
    per-key-point kernels) -- all equal the C restatement bit for bit.  (The library uses the binned form on dense frames
    only; SVI_MATCH_BINNED=2 forces it wherever the geometry allows.)"""
    monkeypatch.setenv("SVI_MATCH_BINNED", "2")
    from oracle import c_oracle as co
    from svi_mapper_b200.synth import stereo_batch_torch
    import torch
    cams = request.getfixturevalue(cams_name)
    W, H = cams[0].width, cams[0].height
    n = 40
    dL, dR = stereo_batch_torch(n, W, H, seed=7300, device=torch.device("cuda", 0))
    Ls, Rs = dL.cpu().numpy(), dR.cpu().numpy()
    Ls[5] = 128                                   # a flat frame: no corner, every bin empty
    masks = np.full_like(Ls, 255)
    masks[7] = 0                                  # a fully masked frame
    masks[9, :, : W // 2] = 0                     # key-points in the right half only: empty and crowded bins side by side
    cfg = co.make_config(cams[0], cams[1], max_corners=max_corners, search_range=search_range)
    ref = co.stereo_frames(cfg, Ls, Rs, masks=masks, n_threads=co.host_threads())
    with StereoFrontend(*cams, max_corners=max_corners, search_range_px=search_range) as fe:
        got = fe.stereo_frames(Ls, Rs, masks)
    n_ok = 0
    for f in range(n):
        r, g_ = co.frame(ref, f), got.frame(f)
        assert len(r["status"]) == len(g_["status"])
        assert (len(r["status"]) == 0) if f in (5, 7) else (len(r["status"]) > 300)
        for k in ("uv_l", "desc_l", "status", "dist", "idx"):
            np.testing.assert_array_equal(g_[k], r[k], err_msg=f"frame {f} {k}")
        ok = r["status"] == 0
        n_ok += int(ok.sum())
        for k in ("uv_r", "desc_r", "xyz"):
            np.testing.assert_array_equal(g_[k][ok], r[k][ok], err_msg=f"frame {f} {k}")
    assert n_ok > 100 * (n - 3)


def test_candidate_overflow_is_reported_on_both_entry_points(kitti_cams):
    """A ctx sized too small for a frame's candidates never truncates silently: the host entry point returns
    SVI_ERR_CAPACITY, the device-resident one reports it through svi_check_overflow, and the condition is cleared."""
    import torch
    from svi_mapper_b200 import SviError
    W, H = kitti_cams[0].width, kitti_cams[0].height
    L, R = stereo_pair(W, H, 11)
    with StereoFrontend(*kitti_cams, max_candidates=1024) as fe:   # ~8000 candidates per textured KITTI-size frame
        with pytest.raises(SviError) as e:
            fe.stereo_frames(L, R)
        assert e.value.code == _lib.SVI_ERR_CAPACITY
        fe.check_overflow()                                        # cleared by the failing call
        dev = torch.device("cuda", 0)
        dL, dR = torch.from_numpy(L).to(dev)[None].contiguous(), torch.from_numpy(R).to(dev)[None].contiguous()
        cap = fe.max_corners
        t = dict(n_kp=torch.full((1,), -1, dtype=torch.int32, device=dev), uv_l=torch.zeros(1, cap, 2, device=dev),
                 uv_r=torch.zeros(1, cap, 2, device=dev), xyz=torch.zeros(1, cap, 3, dtype=torch.float64, device=dev),
                 dl=torch.zeros(1, cap, 32, dtype=torch.uint8, device=dev), dr=torch.zeros(1, cap, 32, dtype=torch.uint8, device=dev),
                 dist=torch.zeros(1, cap, dtype=torch.int32, device=dev), idx=torch.zeros(1, cap, dtype=torch.int32, device=dev),
                 st=torch.zeros(1, cap, dtype=torch.uint8, device=dev))
        r = _lib.StereoResult(cap, t["n_kp"].data_ptr(), None, t["uv_l"].data_ptr(), t["uv_r"].data_ptr(), t["xyz"].data_ptr(),
                              t["dl"].data_ptr(), t["dr"].data_ptr(), t["dist"].data_ptr(), t["idx"].data_ptr(), t["st"].data_ptr())
        fe.stereo_frames_device(dL.data_ptr(), dR.data_ptr(), W, W * H, 1, r, stream=torch.cuda.current_stream().cuda_stream)
        with pytest.raises(SviError) as e:
            fe.check_overflow()
        assert e.value.code == _lib.SVI_ERR_CAPACITY
        assert int(t["n_kp"].cpu()[0]) == 0                        # an overflowing frame yields no key-points at all
        fe.check_overflow()
    with StereoFrontend(*kitti_cams) as fe:                         # default sizing: the same frame is fine
        assert fe.stereo_frames(L, R).n_keypoints[0] > 500
        fe.check_overflow()


def _c_vs_gpu_tracks(got, ref):
    np.testing.assert_array_equal(got["stage"], ref["stage"])
    np.testing.assert_array_equal(got["status"], ref["status"])
    hit = ref["stage"] > 0
    for k in ("uv_l", "uv_r", "xyz", "desc_l", "desc_r"):
        np.testing.assert_array_equal(got[k][hit], ref[k][hit], err_msg=k)


@pytest.mark.parametrize("table", ["alt_random", "alt_adversarial"])
def test_other_pair_tables(vi_cams, kitti_cams, table):
    """Descriptor parity is table-proof: the genuine opencv_contrib pair table (generated_32.i) is not available here
    (DESIGN.md section 2), the shipped one is a stand-in, so swapping the table must be a no-code-change operation.
    libsvi_gpu.so and the C checker are rebuilt around two other tables (uniform random; an adversarial one with all
    offsets at +-24, identical points, repeated and mirrored pairs, odd-only / even-only columns) with the one-command
    recipe `python -m svi_mapper_b200.build --table <txt>`, and the new-landmark path and the whole tracking cascade are
    compared bit for bit -- the pair offsets are template immediates of the match kernel, so this exercises the unrolled
    tests, the gather kernels and the C-ABI end to end."""
    from oracle import c_oracle as co
    from svi_mapper_b200 import build as bld
    alt = dict(bld.build_alt_tables()[table])
    alt["oracle"] = co.build_variant(alt["header"], alt["header"].parent / "libsvi_oracle.so")
    for cams, seed, mc in ((kitti_cams, 0, 1000), (vi_cams, 4000, 1000)):
        W, H = cams[0].width, cams[0].height
        L, R = stereo_pair(W, H, seed)
        cfg = co.make_config(cams[0], cams[1], max_corners=mc)
        ref = co.frame(co.stereo_frames(cfg, L, R, native=alt["oracle"]), 0)
        base = co.frame(co.stereo_frames(cfg, L, R), 0)
        assert not np.array_equal(ref["desc_l"], base["desc_l"])            # it really is another table
        with StereoFrontend(*cams, lib_path=alt["lib"], max_corners=mc) as fe:
            got = fe.add_new_landmarks(L, R)
            for k in ("uv_l", "desc_l", "status", "dist", "idx"):
                np.testing.assert_array_equal(got[k], ref[k], err_msg=k)
            good = ref["status"] == 0
            assert good.sum() > 300
            for k in ("uv_r", "desc_r", "xyz"):
                np.testing.assert_array_equal(got[k][good], ref[k][good], err_msg=k)
            if cams is not vi_cams:
                continue
            # tracking with the same table: stage 1 (unchanged pair), stage 2 (shifted pair), stage 3 (moved camera)
            ok = np.nonzero(good)[0]
            disp = (ref["uv_l"][ok, 0] - ref["uv_r"][ok, 0]).astype(np.float32)
            args = (ref["xyz"][ok], ref["desc_l"][ok], ref["desc_r"][ok], disp, 7.0)
            kw = dict(uv_reference_left=ref["uv_l"][ok], desc_reference_left=ref["desc_l"][ok], T_left_to_world_at_detection=np.eye(4))
            T = np.eye(4)
            T[:3, 3] = (0.015, 0.01, 0.0)
            seen = np.zeros(6, np.int64)
            for a, b, pose, scaling in ((L, R, np.eye(4), 1.0), (np.roll(L, (2, 3), (0, 1)), np.roll(R, (2, 3), (0, 1)), np.eye(4), 1.0), (L, R, T, 1.5)):
                want = co.track_landmarks(cfg, a, b, pose, *args, scaling, native=alt["oracle"], n_threads=co.host_threads(), **kw)
                _c_vs_gpu_tracks(fe.track_landmarks(a, b, pose, *args, scaling, **kw), want)
                seen += np.bincount(want["stage"], minlength=6)
            assert seen[1] > 200 and seen[3] > 100 and seen[5] > 20, seen


def test_track_cascade_full_landmark_set_vs_c_port(vi_cams):
    """The whole cascade on every landmark of a vi_sensor frame (the C3 shape: thousands of landmarks per frame) against
    the C port of trackManual -- the numpy restatement covers a few hundred landmarks per test, the C port all of them."""
    from oracle import c_oracle as co
    W, H = vi_cams[0].width, vi_cams[0].height
    L, R = stereo_pair(W, H, 4001)
    cfg = co.make_config(vi_cams[0], vi_cams[1], max_corners=3000)
    ref0 = co.frame(co.stereo_frames(cfg, L, R), 0)
    ok = np.nonzero(ref0["status"] == 0)[0]
    assert len(ok) > 1500
    disp = (ref0["uv_l"][ok, 0] - ref0["uv_r"][ok, 0]).astype(np.float32)
    args = (ref0["xyz"][ok], ref0["desc_l"][ok], ref0["desc_r"][ok], disp, 7.0)
    kw = dict(uv_reference_left=ref0["uv_l"][ok], desc_reference_left=ref0["desc_l"][ok], T_left_to_world_at_detection=np.eye(4))
    L1, R1 = np.roll(L, (1, 2), (0, 1)), np.roll(R, (1, 2), (0, 1))
    seen = np.zeros(6, np.int64)
    with StereoFrontend(*vi_cams, max_corners=3000) as fe:
        for a, b, t, scaling in ((L, R, (0.0, 0.0, 0.0), 1.0), (L1, R1, (0.0, 0.0, 0.0), 1.0), (L1, R1, (0.02, 0.01, 0.03), 2.0), (L, R, (0.01, 0.0, 0.02), 1.3)):
            T = np.eye(4)
            T[:3, 3] = t
            want = co.track_landmarks(cfg, a, b, T, *args, scaling, n_threads=co.host_threads(), **kw)
            _c_vs_gpu_tracks(fe.track_landmarks(a, b, T, *args, scaling, **kw), want)
            seen += np.bincount(want["stage"], minlength=6)
    assert seen[1] > 1000 and seen[3] > 500 and seen[5] > 50, seen


def test_detection_mask_on_gpu_matches_cv2_circle(kitti_cams, calib_dir, tmp_path):
    """getMaskActiveLandmarks (CFundamentalMatcher.cpp:2043-2073) stamped by the GPU stencil kernel == the cv2.circle
    golden plane (tests/golden/stereo_320x240.npz: 60 centres incl. discs cut by every image edge), through the C-ABI
    and through the C++ facade (facade_demo --mask); addNewLandmarks with device-built mask == with an uploaded plane."""
    import pathlib
    import subprocess
    from svi_mapper_b200.calib import PinholeCamera
    gold = np.load(pathlib.Path(__file__).resolve().parent / "golden" / "stereo_320x240.npz")
    small = [PinholeCamera(c.label, 320, 240, c.P, c.K, c.focal_length_m, c.distortion, c.rectification) for c in kitti_cams]
    with StereoFrontend(*small) as fe:
        np.testing.assert_array_equal(fe.mask_active_landmarks(gold["mask_centres"]), gold["mask"])
        assert (fe.mask_active_landmarks(np.zeros((0, 2), np.float32)) == 255).all()
        far = np.array([[1e9, 5.0], [np.nan, 3.0], [-1e12, -1e12], [160.4, 120.6]], np.float32)   # projections at infinity draw nothing
        np.testing.assert_array_equal(fe.mask_active_landmarks(far), o.mask_active_landmarks(320, 240, [(160.4, 120.6)]))
        # new landmarks under the mask: centres on the device == the same plane uploaded by the caller
        a = fe.add_new_landmarks(gold["left"], gold["right"], mask=gold["mask"])
        b = fe.add_new_landmarks(gold["left"], gold["right"], mask_centres=gold["mask_centres"])
        c = fe.add_new_landmarks(gold["left"], gold["right"], mask_centres=np.zeros((0, 2), np.float32))
        d = fe.add_new_landmarks(gold["left"], gold["right"])
        for k in a:
            np.testing.assert_array_equal(a[k], b[k], err_msg=k)
            np.testing.assert_array_equal(c[k], d[k], err_msg=k)
        assert len(a["status"]) > 50 and len(a["status"]) != len(d["status"])
    with StereoFrontend(*small, max_corners=300) as fe:       # the golden frame was produced with maxCorners 300
        np.testing.assert_array_equal(fe.add_new_landmarks(gold["left"], gold["right"], mask_centres=np.zeros((0, 2), np.float32))["uv_l"], gold["frame_uv_l"])
    # the C++ facade draws the same plane (one implementation: the kernel)
    exe = pathlib.Path(__file__).resolve().parents[1] / "svi_mapper_b200" / "host" / "facade_demo"
    for side in ("left", "right"):
        txt = (calib_dir / f"kitti_00_{side}.txt").read_text().replace("uWidthPixels 1241", "uWidthPixels 320").replace("uHeightPixels 376", "uHeightPixels 240")
        (tmp_path / f"{side}.txt").write_text(txt)
    np.savetxt(tmp_path / "centres.txt", gold["mask_centres"], fmt="%.9g")
    r = subprocess.run([str(exe), "--mask", str(tmp_path / "left.txt"), str(tmp_path / "right.txt"), str(tmp_path / "centres.txt"), str(tmp_path / "mask.raw")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    np.testing.assert_array_equal(np.fromfile(tmp_path / "mask.raw", np.uint8).reshape(240, 320), gold["mask"])


class _CpuBackend:
    """SequenceTracker backend over the C restatement (checker side of the sequence tests)."""

    def __init__(self, cfg, threads):
        from oracle import c_oracle as co
        self.co, self.cfg, self.threads = co, cfg, threads

    def track(self, L, R, T, s, scaling, size):
        return self.co.track_landmarks(self.cfg, L, R, T, s["xyz_w"], s["last_desc_l"], s["last_desc_r"], s["last_disp"], size, scaling,
                                       uv_reference_left=s["uv_ref"], desc_reference_left=s["ref_desc_l"],
                                       T_left_to_world_at_detection=s["T_det"], n_threads=self.threads)

    def add_new(self, L, R, centres):
        m = o.mask_active_landmarks(L.shape[1], L.shape[0], centres)[None] if len(centres) else None
        return self.co.frame(self.co.stereo_frames(self.cfg, L, R, masks=m), 0)


def test_sequence_tracking_in_lockstep_with_cpu_port(vi_cams):
    """BASELINE configs[2] in small: a rendered vi_sensor sequence (static world, smooth trajectory) run through the
    per-frame loop of CTrackerGT::_trackLandmarks -- trackManual on every active landmark, visibility / failure
    bookkeeping, retirement, re-detection under the device-built mask -- once over the GPU front-end and once over the C
    port, with the landmark state fed forward.  Every frame's tracking result, new-landmark result and bookkeeping
    record must be identical (so the two sequences never diverge), and all three stages must occur."""
    from oracle import c_oracle as co
    from svi_mapper_b200.sequence import GpuBackend, SequenceTracker, render_sequence
    n = 14
    L, R, T = render_sequence(vi_cams[0], vi_cams[1], n)
    cfg = co.make_config(vi_cams[0], vi_cams[1], max_corners=1000)
    total = np.zeros(6, np.int64)
    with StereoFrontend(*vi_cams) as fe:
        g, c = SequenceTracker(GpuBackend(fe), vi_cams[0]), SequenceTracker(_CpuBackend(cfg, co.host_threads()), vi_cams[0])
        for t in range(n):
            rg, rc = g.process(L[t], R[t], T[t]), c.process(L[t], R[t], T[t])
            assert rg == rc, (t, rg, rc)
            if rg["tracked"]:
                _c_vs_gpu_tracks(g.last_track, c.last_track)
            if rg["new"]:
                for k in ("uv_l", "desc_l", "status"):
                    np.testing.assert_array_equal(g.last_new[k], c.last_new[k], err_msg=k)
                ok = c.last_new["status"] == 0
                for k in ("uv_r", "desc_r", "xyz"):
                    np.testing.assert_array_equal(g.last_new[k][ok], c.last_new[k][ok], err_msg=k)
            for k in g.s:
                np.testing.assert_array_equal(g.s[k], c.s[k], err_msg=k)
            total += np.asarray(rg["stages"])
    assert g.n_active > 1500 and total[1] > 5000 and total[3] + total[4] > 100 and total[5] > 100, (g.n_active, total)
    assert sum(r["new"] > 0 for r in g.log) >= 3          # the re-detection trigger of CTrackerGT.cpp:305 fired repeatedly


def _same_stereo(a, b, n):
    np.testing.assert_array_equal(a.n_keypoints, b.n_keypoints)
    np.testing.assert_array_equal(a.n_detected, b.n_detected)
    for f in range(n):
        fa, fb = a.frame(f), b.frame(f)
        for k in ("uv_l", "desc_l", "status", "dist", "idx"):
            assert fa[k].tobytes() == fb[k].tobytes(), (f, k)
        ok = fa["status"] == 0
        for k in ("uv_r", "desc_r", "xyz"):
            assert fa[k][ok].tobytes() == fb[k][ok].tobytes(), (f, k)


def test_multi_gpu_driver_equals_single_device(kitti1112_cams):
    """svi_multi (one host thread + one svi_ctx per device, contiguous frame ranges, disjoint output slices) returns
    byte-for-byte what one device returns.  On a box with >= 2 GPUs the ranges really run on different
    devices; with one GPU the same driver runs two contexts of device 0 side by side (same code path, same threads)."""
    import torch
    W, H = kitti1112_cams[0].width, kitti1112_cams[0].height
    n = 37                                               # not divisible by 2, 3 or 4: ragged ranges
    pairs = [stereo_pair(W, H, 2000 + i) for i in range(6)]
    L = np.stack([pairs[i % 6][0] for i in range(n)])
    R = np.stack([pairs[i % 6][1] for i in range(n)])
    for i in range(n):                                   # make every frame distinct
        L[i, :, : 8 + i] = L[i, :, : 8 + i][:, ::-1]
    masks = np.full((n, H, W), 255, np.uint8)
    masks[::3, 100:200, 300:700] = 0
    from svi_mapper_b200 import MultiFrontend
    with StereoFrontend(*kitti1112_cams, chunk_frames=8) as fe:
        one = fe.stereo_frames(L, R, masks)
    ndev = torch.cuda.device_count()
    for devices in ([0, 1 % ndev], [0, 1 % ndev, 2 % ndev], [d % ndev for d in range(4)], [0]):
        with MultiFrontend(*kitti1112_cams, devices=devices, chunk_frames=8) as mf:
            assert [mf.frame_range(n, g) for g in range(len(devices))] == [((g * n) // len(devices), ((g + 1) * n) // len(devices)) for g in range(len(devices))]
            _same_stereo(mf.stereo_frames(L, R, masks), one, n)
            _same_stereo(mf.stereo_frames(L[:1], R[:1], masks[:1]), fe_one_frame(kitti1112_cams, L[:1], R[:1], masks[:1]), 1)   # fewer frames than devices


def fe_one_frame(cams, L, R, M):
    with StereoFrontend(*cams, chunk_frames=8) as fe:
        return fe.stereo_frames(L, R, M)


def test_cpp_tracker_sequence_matches_cpu_restatement(vi_cams, calib_dir, tmp_path):
    """Sequence-level parity of the host orchestration (CTrackerGT::_trackLandmarks, src/core/CTrackerGT.cpp:137-380): the
    C++ layer (facade_demo --sequence: CTrackerGT::process over the GPU front-end -- trackManual, CLandmark::addMeasurement /
    optimize, retirement, masked re-detection) against oracle/tracker_np.py running the same loop over the CPU restatement,
    frame by frame: counters, every visible landmark's id / measurement, and every active landmark's state (optimised
    position, optimal / visible flags, optimisation and failure counters, number of measurements)."""
    import pathlib
    import subprocess
    from oracle import c_oracle as co
    from oracle import tracker_np as tn
    from svi_mapper_b200.sequence import render_sequence
    n, mc = 12, 150
    L, R, T = render_sequence(vi_cams[0], vi_cams[1], n, 4100)
    L.tofile(tmp_path / "L.raw")
    R.tofile(tmp_path / "R.raw")
    rel, rot = [], []
    for t in range(n):
        M = np.eye(4) if t == 0 else T[t] @ np.linalg.inv(T[t - 1])
        rel.append(M)
        rot.append(float(np.arccos(np.clip((np.trace(M[:3, :3]) - 1.0) / 2.0, -1.0, 1.0))))
    with open(tmp_path / "motions.txt", "w") as f:
        for M, a in zip(rel, rot):
            f.write(" ".join(repr(float(v)) for v in M[:3].reshape(-1)) + " " + repr(a) + "\n")
    exe = pathlib.Path(__file__).resolve().parents[1] / "svi_mapper_b200" / "host" / "facade_demo"
    out = tmp_path / "seq.txt"
    r = subprocess.run([str(exe), "--sequence", str(calib_dir / "vi_sensor_left.txt"), str(calib_dir / "vi_sensor_right.txt"), str(n),
                        str(tmp_path / "L.raw"), str(tmp_path / "R.raw"), str(tmp_path / "motions.txt"), str(mc), str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    frames, cur = [], None
    for line in out.read_text().splitlines():
        w = line.split()
        if w[0] == "F":
            cur = dict(head=dict(zip(w[2::2], map(int, w[3::2]))), M=[], A=[])
            frames.append(cur)
        else:
            cur[w[0]].append(w[1:])
    assert len(frames) == n
    trk = tn.TrackerGT(vi_cams[0], vi_cams[1], co.make_config(vi_cams[0], vi_cams[1], max_corners=mc), threads=co.host_threads())
    optimised = retired = 0
    for t in range(n):
        before = {lm.uid for lm in trk.active()}
        trk.process(L[t], R[t], rel[t], rot[t])
        fr = frames[t]
        act = trk.active()
        assert fr["head"] == dict(VISIBLE=trk.visible_last, ACTIVE=len(act), S1=trk.tracks[0], S2=trk.tracks[1], S3=trk.tracks[2], DETECTIONS=trk.detections), (t, fr["head"])
        assert [int(m[0]) for m in fr["M"]] == [lm.uid for lm in trk.visible], t
        for m, lm in zip(fr["M"], trk.visible):
            meas = lm.measurements[-1]
            assert tuple(np.float32(v) for v in m[1:5]) == (meas[2][0], meas[2][1], meas[3][0], meas[3][1]), (t, lm.uid)
            np.testing.assert_allclose([float(v) for v in m[5:8]], lm.last_xyz_left, rtol=1e-9, atol=0)
        assert [int(a[0]) for a in fr["A"]] == [lm.uid for lm in act], t
        for a, lm in zip(fr["A"], act):
            assert [int(v) for v in a[4:10]] == [int(lm.optimal), int(lm.visible), lm.opt_success, lm.opt_failed, lm.failed, len(lm.measurements)], (t, lm.uid, a)
            np.testing.assert_allclose([float(v) for v in a[1:4]], lm.xyz_opt, rtol=1e-6, atol=1e-9)
        optimised += sum(lm.opt_success for lm in act if len(lm.measurements) == 6)
        retired += len(before - {lm.uid for lm in act})
    assert trk.detections >= 3 and optimised > 50 and sum(trk.tracks) > 100, (trk.detections, optimised, retired)


def test_optimize_landmarks_batch_matches_oracle_and_cpp(vi_cams, calib_dir, tmp_path):
    """svi_optimize_landmarks (CLandmark::optimize for a whole set of landmarks in one launch, reference
    src/types/CLandmark.cpp:281-296, :447-581) against the numpy restatement landmark by landmark -- position within 1e-7 m,
    the optimal / converged / failed verdict identical -- on 400 landmarks with 1 ... 45 measurements over 50 poses: clean
    tracks, noisy ones, outlier-dominated ones, too few measurements.  Then the C++ host layer run over the same rendered
    sequence twice, once through this call and once through its own CPU loop (SVI_HOST_OPTIMIZE=cpu): every printed number
    of every frame is identical, i.e. the GPU arithmetic equals the host's bit for bit."""
    import pathlib
    import subprocess
    import oracle.frontend_np as o
    from svi_mapper_b200.sequence import render_sequence
    P_l, P_r = np.asarray(vi_cams[0].P, np.float64).reshape(3, 4), np.asarray(vi_cams[1].P, np.float64).reshape(3, 4)
    rng = np.random.default_rng(11)
    n_poses, n = 50, 400
    Tw = []
    for k in range(n_poses):
        T = np.eye(4)
        a = 0.004 * k
        T[:3, :3] = [[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]]
        T[:3, 3] = [0.03 * k, -0.01 * k, -0.04 * k]
        Tw.append(T)
    PL, PR = np.stack([P_l @ T for T in Tw]), np.stack([P_r @ T for T in Tw])
    guess, first, pose, uvl, uvr, per_lm = [], [0], [], [], [], []
    for i in range(n):
        truth = np.array([rng.uniform(-2, 2), rng.uniform(-1, 1), rng.uniform(3, 25)])
        kind = i % 4                                   # 0 clean, 1 noisy, 2 outlier-dominated, 3 short
        cnt = int(rng.integers(1, 6)) if kind == 3 else int(rng.integers(6, 46))
        start = int(rng.integers(0, n_poses - cnt + 1))
        ms = []
        for k in range(start, start + cnt):
            a, b = PL[k] @ np.append(truth, 1), PR[k] @ np.append(truth, 1)
            noise = 0.2 if kind != 1 else 1.5
            l = np.float32([a[0] / a[2], a[1] / a[2]]) + np.float32(rng.normal(0, noise, 2))
            r = np.float32([b[0] / b[2], l[1]]) + np.float32([rng.normal(0, noise), 0])
            if kind == 2 and (k - start) % 3:
                l += np.float32(rng.uniform(15, 50, 2))
            ms.append((PL[k], PR[k], l, r))
            pose.append(k); uvl.append(l); uvr.append(r)
        first.append(len(pose))
        x0 = truth + np.array([0.1, -0.05, 0.6]) * rng.uniform(0.2, 1.5)
        guess.append(x0)
        per_lm.append((x0, ms))
    with StereoFrontend(*vi_cams) as fe:
        got = fe.optimize_landmarks(np.array(guess), first, pose, np.array(uvl), np.array(uvr), PL, PR)
        empty = fe.optimize_landmarks(np.zeros((0, 3)), [0], [], np.zeros((0, 2)), np.zeros((0, 2)), PL, PR)
        assert len(empty["outcome"]) == 0
        with pytest.raises(SviError):
            fe.optimize_landmarks(np.array(guess[:2]), [0, 1, 2], [0, n_poses], np.array(uvl[:2]), np.array(uvr[:2]), PL, PR)   # pose row out of range
    seen = set()
    for i, (x0, ms) in enumerate(per_lm):
        ref = o.optimize_landmark(x0, ms)
        oc = int(got["outcome"][i])
        seen.add(oc)
        if len(ms) <= 5:
            assert oc == 0 and ref["optimal"] and np.array_equal(got["xyz"][i], x0)
            continue
        assert (oc in (1, 2), oc in (3, 4), oc == 2) == (ref["success"] == 1, ref["failed"] == 1, bool(ref["optimal"])), (i, oc, ref)
        np.testing.assert_allclose(got["xyz"][i], ref["xyz"], rtol=0, atol=1e-7, err_msg=str(i))
    assert {0, 2, 3} <= seen                            # skipped, optimal and rejected all occur

    # extreme magnitudes: far points, a guess a million times too far, and the run-away landmark captured from
    # the C3 sequence (position 1e83 m, exhausts the 1000 iterations) -- bit for bit against the host's IEEE arithmetic
    exe = pathlib.Path(__file__).resolve().parents[1] / "svi_mapper_b200" / "host" / "facade_demo"
    golden = pathlib.Path(__file__).resolve().parent / "golden" / "landmark_runaway.txt"
    cases = []
    rows = [ln for ln in golden.read_text().splitlines() if ln and not ln.startswith("#")]
    cases.append((np.array([float(v) for v in rows[0].split()]),
                  [(np.array(v[:12]), np.array(v[12:24]), np.float32(v[24:26]), np.float32(v[26:28])) for v in ([float(t) for t in r.split()] for r in rows[1:])]))
    for z, scale in ((1e3, 1.3), (1e5, 0.7), (1e8, 2.0), (1e12, 1.1), (40.0, 1e6)):
        truth = np.array([0.3 * z, -0.1 * z, z])
        ms = []
        for k in range(9):
            a, b = PL[3 * k] @ np.append(truth, 1), PR[3 * k] @ np.append(truth, 1)
            l = np.float32([a[0] / a[2], a[1] / a[2]]) + np.float32(rng.normal(0, 0.3, 2))
            ms.append((PL[3 * k], PR[3 * k], l, np.float32([b[0] / b[2], l[1]]) + np.float32([rng.normal(0, 0.3), 0])))
        cases.append((truth * scale, ms))
    with StereoFrontend(*vi_cams) as fe:
        for ci, (x0, ms) in enumerate(cases):
            g1 = fe.optimize_landmarks(x0[None], [0, len(ms)], np.arange(len(ms)), np.array([m[2] for m in ms]), np.array([m[3] for m in ms]),
                                       np.stack([np.asarray(m[0]).reshape(12) for m in ms]), np.stack([np.asarray(m[1]).reshape(12) for m in ms]))
            f = tmp_path / f"lm_{ci}.txt"
            f.write_text("\n".join([" ".join(repr(float(v)) for v in x0)] +
                                    [" ".join(repr(float(v)) for v in list(np.asarray(m[0]).ravel()) + list(np.asarray(m[1]).ravel()) + [m[2][0], m[2][1], m[3][0], m[3][1]]) for m in ms]) + "\n")
            r = subprocess.run([str(exe), "--landmark", str(f)], capture_output=True, text=True)
            assert r.returncode == 0, r.stderr
            v = r.stdout.split()
            oc = int(g1["outcome"][0])
            assert (int(v[3]), int(v[4]), int(v[5])) == (int(oc == 2), int(oc in (1, 2)), int(oc in (3, 4))), (ci, oc, v)
            assert [repr(float(t)) for t in g1["xyz"][0]] == [repr(float(t)) for t in v[:3]], (ci, g1["xyz"][0], v[:3])
        assert int(fe.optimize_landmarks(cases[0][0][None], [0, len(cases[0][1])], np.arange(len(cases[0][1])), np.array([m[2] for m in cases[0][1]]),
                                         np.array([m[3] for m in cases[0][1]]), np.stack([m[0] for m in cases[0][1]]), np.stack([m[1] for m in cases[0][1]]))["iterations"][0]) == 1000

    # the C++ tracker over a rendered sequence: library call vs the host's own CPU loop
    nf, mc = 14, 200
    L, R, T = render_sequence(vi_cams[0], vi_cams[1], nf, 4200)
    L.tofile(tmp_path / "L.raw")
    R.tofile(tmp_path / "R.raw")
    with open(tmp_path / "motions.txt", "w") as f:
        for t in range(nf):
            M = np.eye(4) if t == 0 else T[t] @ np.linalg.inv(T[t - 1])
            f.write(" ".join(repr(float(v)) for v in M[:3].reshape(-1)) + " " + repr(float(np.arccos(np.clip((np.trace(M[:3, :3]) - 1.0) / 2.0, -1.0, 1.0)))) + "\n")
    exe = pathlib.Path(__file__).resolve().parents[1] / "svi_mapper_b200" / "host" / "facade_demo"
    outs = []
    for mode in ("gpu", "cpu"):
        out = tmp_path / f"seq_{mode}.txt"
        import os
        r = subprocess.run([str(exe), "--sequence", str(calib_dir / "vi_sensor_left.txt"), str(calib_dir / "vi_sensor_right.txt"), str(nf),
                            str(tmp_path / "L.raw"), str(tmp_path / "R.raw"), str(tmp_path / "motions.txt"), str(mc), str(out)],
                           capture_output=True, text=True, env=dict(os.environ, SVI_HOST_OPTIMIZE=mode))
        assert r.returncode == 0, r.stderr
        outs.append(out.read_text())
    assert outs[0] == outs[1] and outs[0].count("\nA ") > 500
    assert any(int(line.split()[7]) > 0 for line in outs[0].splitlines() if line.startswith("A "))   # some landmark was optimised successfully


def test_bounds_checked_build():
    """compute-sanitizer (memcheck / racecheck / synccheck) is closed on the GPU pool this repository is developed on, so the
    kernels carry their own assertions: in the -DSVI_BOUNDS_CHECK build every index that is computed at run time and goes into
    shared memory, a candidate / item list or an output array is checked, and a violation turns the next call into an error.
    The parity tests with the widest coverage of those indices run once against that build: whole frames incl. the window /
    ROI detector, both selection variants (shared / global), the long scan lines, masks, the tracking cascade, the binned matcher."""
    import os
    import pathlib
    import subprocess
    import sys
    from svi_mapper_b200 import build as bld
    lib = bld.build_checked()
    k = ("test_stereo_frame_parity or test_track_manual_stage2_window_search or test_stress_frame_global_select_and_long_scanlines "
         "or test_stereo_batch_chunks_and_masks or test_track_manual_stage3_epipolar or test_edge_cases_empty_flat_masked_padded "
         "or test_binned_matcher_geometries or test_optimize_landmarks_batch_matches_oracle_and_cpp")
    env = dict(os.environ, SVI_GPU_LIB=str(lib))
    r = subprocess.run([sys.executable, "-m", "pytest", str(pathlib.Path(__file__)), "-x", "-q", "-k", k], env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout and "failed" not in r.stdout
