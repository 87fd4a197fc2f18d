"""TEST INFRASTRUCTURE — writes tests/golden/*.npz (run in the build container only; needs cv2).

    python -m oracle.pyref.make_goldens [orb|match|triang|fuse|loop|junctions|line|linematch|linefuse|linetriangnew|pose|pose_lil|undistort|planes|lines3d|all]

Each fixture stores the seeded input bytes and the outputs of the cv2-primitive
restatement of the reference (oracle/pyref), so that the tests never need cv2 or
/root/reference at run time.
"""
from __future__ import annotations

import os
import sys

import numpy as np

from oracle.pyref import orb_cv2
from psl_slam_b200 import synth

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "tests", "golden")


def orb_cases():
    g, _, _ = synth.sequence(1, 1)
    yield "orb_vga_seed1", g[0], dict(nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7)
    rng = np.random.default_rng(11)
    yield "orb_noise_200x160", rng.integers(0, 256, (160, 200), dtype=np.uint8), dict(
        nfeatures=500, scale_factor=1.2, nlevels=4, ini_th=20, min_th=7)
    poster = synth.make_poster(12, 1024, 900)
    rgb, _ = synth.render(poster, synth.trajectory(1, 12)[0], 752, 300, noise_seed=5)
    yield "orb_wide_752x300", synth.rgb_to_gray(rgb), dict(nfeatures=800, scale_factor=1.2, nlevels=6, ini_th=20,
                                                          min_th=7)
    yield "orb_lowtex_320x240", synth.make_lowtex(3, 320, 240), dict(nfeatures=300, scale_factor=1.2, nlevels=5,
                                                                     ini_th=20, min_th=7)
    poster = synth.make_poster(5, 2048)
    rgb, _ = synth.render(poster, synth.trajectory(1, 5)[0], 960, 540, noise_seed=9)
    yield "orb_qhd_960x540_12lvl", synth.rgb_to_gray(rgb), dict(nfeatures=4000, scale_factor=1.2, nlevels=12,
                                                               ini_th=20, min_th=7)
    yield "orb_flat_160x120", np.full((120, 160), 77, np.uint8), dict(nfeatures=100, scale_factor=1.2, nlevels=3,
                                                                     ini_th=20, min_th=7)


def make_orb():
    for name, img, kw in orb_cases():
        P = orb_cv2.OrbParams(**kw)
        st = {}
        kps, octv, desc = orb_cv2.orb_extract(img, P, st)
        arrays = dict(image=img, kps=kps, octave=octv, desc=desc, params=np.array(
            [kw["nfeatures"], kw["nlevels"], kw["ini_th"], kw["min_th"]], np.int32),
            scale_factor=np.float32(kw["scale_factor"]))
        for l in range(P.nlevels):
            arrays[f"cands_{l}"] = st["cands"][l]
            arrays[f"sel_{l}"] = st["selected"][l]
        # one mid-pyramid level image + its blur pin the resize chain and the blur
        l = min(2, P.nlevels - 1)
        arrays["level_idx"] = np.int32(l)
        arrays["level_img"] = st["level_img"][l]
        if l in st.get("blur", {}):
            arrays["level_blur"] = st["blur"][l]
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrays)
        print(name, img.shape, "n =", len(kps), "per-level", np.bincount(octv, minlength=P.nlevels).tolist())


def make_match():
    """cfg-2 shaped fixtures: consecutive RGB-D frames, projection queries, BoW groups, knn2."""
    from oracle import orc
    from oracle.pyref import frame_py, match_py
    K = synth.ICL
    gray, depth, T = synth.sequence(2, 3)
    P = orb_cv2.OrbParams()
    ext = [orc.orb_extract(g) for g in gray]  # bit-identical to orb_cv2 (tests/test_oracle_orb.py)
    bounds = (np.float32(0), np.float32(0), np.float32(640), np.float32(480))  # zero distortion: Frame.cc:155-158
    scale = np.array(P.scale, np.float32)
    views, world = [], []
    for i in range(3):
        kps, desc = ext[i]
        xy = np.stack([kps["x"], kps["y"]], 1)
        ur, dep = frame_py.stereo_from_rgbd(xy, frame_py.depth_to_float(depth[i]), K["bf"])
        Twc = np.linalg.inv(T[i])
        world.append(frame_py.unproject(xy, dep, K, Twc))
        views.append((kps, desc, ur))
    rng = np.random.default_rng(22)
    for pair, (a, b) in enumerate([(0, 1), (1, 2)]):
        kl, dl, _ = views[a]
        kc, dc, urc = views[b]
        n = len(kl)
        outlier = rng.random(n) < 0.05
        has_mp = rng.random(n) < 0.9
        claims = rng.random(n) < 0.8  # temporal points (Observations()==0) do not claim
        # pose prior = true pose perturbed a little (the motion model is never exact)
        Tcur = T[b].copy()
        Tcur[:3, 3] += rng.normal(0, 0.002, 3)
        q = frame_py.projection_queries(world[a], kl["octave"], kl["angle"], has_mp & ~outlier, claims, Tcur, T[a], K,
                                        scale, 15.0 if pair == 0 else 30.0, bounds)
        kun = np.stack([kc["x"], kc["y"], kc["size"], kc["angle"], kc["response"]], 1).astype(np.float32)
        fv = match_py.FrameView(kun, kc["octave"], urc, dc, *bounds)
        assign, nm = match_py.search_by_projection(fv, q, dl, None, 0, 100, 0.9, True)
        # mode 1: same queries read as local-map points (levels pred-1..pred), some keypoints pre-claimed
        q1 = q.copy()
        q1["min_level"] = kl["octave"] - 1
        q1["max_level"] = kl["octave"]
        pre = rng.random(len(dc)) < 0.3
        assign1, nm1 = match_py.search_by_projection(fv, q1, dl, pre, 1, 100, 0.8, False)
        # BoW: fake vocabulary = 6 descriptor bits -> 64 nodes, a few nodes missing on either side
        def fvec(desc, drop):
            node = (desc[:, 0] & 0x3F).astype(np.uint32)
            d = {}
            for i, nd in enumerate(node):
                if nd % 7 != drop:
                    d.setdefault(int(nd), []).append(i)
            return d
        kf_fv, f_fv = fvec(dl, 3), fvec(dc, 5)
        match, nmb = match_py.search_by_bow(dl, kl["angle"], has_mp, kf_fv, dc, kc["angle"], f_fv, 0.7, 50, True)
        idx2, dist2 = match_py.knn2_cv2(dl[:200], dc[:200])
        def csr(d):
            ids = sorted(d)
            offs = np.cumsum([0] + [len(d[k]) for k in ids]).astype(np.int32)
            return np.array(ids, np.uint32), offs, np.array([i for k in ids for i in d[k]], np.uint32)
        kcsr, fcsr = csr(kf_fv), csr(f_fv)
        np.savez_compressed(os.path.join(OUT, f"match_pair{pair}.npz"), kps_cur=kc, desc_cur=dc, u_right_cur=urc,
                            desc_last=dl, angle_last=kl["angle"], bounds=np.array(bounds, np.float32), queries=q,
                            assign0=assign, nmatches0=np.int32(nm), queries1=q1, claimed1=pre.astype(np.uint8),
                            assign1=assign1, nmatches1=np.int32(nm1), kf_valid=has_mp.astype(np.uint8),
                            kf_nodes=kcsr[0], kf_offs=kcsr[1], kf_idx=kcsr[2], f_nodes=fcsr[0], f_offs=fcsr[1],
                            f_idx=fcsr[2], bow_match=match, bow_nmatches=np.int32(nmb), knn_idx=idx2, knn_dist=dist2)
        print(f"match_pair{pair}: queries {int((q['flags'] & 1).sum())} proj-matches {nm} / mode1 {nm1} / bow {nmb}")
        # relocalization overload (ORBmatcher.cc:1472-1599): the same points read as pKF's MapPoints, window
        # pred-1..pred+1 (pred = the octave they were seen at, jittered), some keypoints already hold a MapPoint
        rng2 = np.random.default_rng(500 + pair)
        q2 = q.copy()
        pred = np.clip(kl["octave"] + rng2.integers(-1, 2, n), 0, P.nlevels - 1)
        q2["min_level"], q2["max_level"] = pred - 1, pred + 1
        q2["radius"] = (np.float32(10.0 if pair == 0 else 20.0) * scale[pred]).astype(np.float32)
        q2["flags"] = (has_mp & (rng2.random(n) < 0.85)).astype(np.uint32)      # not bad, not in sAlreadyFound
        held = rng2.random(len(dc)) < 0.25
        for orb_dist, tag in ((100, "a"), (64, "b")):
            ar, nr = match_py.search_by_projection_keyframe(fv, q2, dl, held, orb_dist, True)
            np.savez_compressed(os.path.join(OUT, f"reloc_pair{pair}{tag}.npz"), kps_cur=kc, desc_cur=dc,
                                bounds=np.array(bounds, np.float32), queries=q2, desc_kf=dl,
                                held=held.astype(np.uint8), orb_dist=np.int32(orb_dist), assign=ar, nmatches=np.int32(nr))
            print(f"reloc_pair{pair}{tag}: queries {int((q2['flags'] & 1).sum())} matches {nr}")


def make_triangulation():
    """N1 fixture: SearchForTriangulation between two keyframes of the synthetic sequence (fake 64-node vocabulary,
    F12 from the true poses as LocalMapping::ComputeF12 does)."""
    from oracle import orc
    from oracle.pyref import frame_py, match_py
    K = synth.ICL
    gray, depth, T = synth.sequence(3, 6)
    P = orb_cv2.OrbParams()
    scale = np.array(P.scale, np.float32)
    sigma2 = (scale * scale).astype(np.float32)
    rng = np.random.default_rng(31)
    for case, (i1, i2, only_stereo) in enumerate([(5, 0, False), (4, 1, True)]):
        views = []
        for i in (i1, i2):
            kps, desc = orc.orb_extract(gray[i])
            xy = np.stack([kps["x"], kps["y"]], 1)
            ur, _ = frame_py.stereo_from_rgbd(xy, frame_py.depth_to_float(depth[i]), K["bf"])
            ur = np.where(rng.random(len(ur)) < 0.5, ur, -1.0).astype(np.float32)  # half the points "monocular"
            has_mp = rng.random(len(kps)) < 0.4
            node = (desc[:, 0] & 0x3F).astype(np.uint32)
            fv = {}
            for k, nd in enumerate(node):
                if nd % 7 != (3 if i == i1 else 5):
                    fv.setdefault(int(nd), []).append(k)
            views.append((kps, ur, desc, has_mp, fv))
        # ComputeF12 (LocalMapping.cc): F12 = K1^-T [t12]x R12 K2^-1 in float32
        T1, T2 = T[i1].astype(np.float32), T[i2].astype(np.float32)
        R1w, t1w, R2w, t2w = T1[:3, :3], T1[:3, 3], T2[:3, :3], T2[:3, 3]
        R12 = (R1w @ R2w.T).astype(np.float32)
        t12 = (-R1w @ R2w.T @ t2w + t1w).astype(np.float32)
        tx = np.array([[0, -t12[2], t12[1]], [t12[2], 0, -t12[0]], [-t12[1], t12[0], 0]], np.float32)
        Km = np.array([[K["fx"], 0, K["cx"]], [0, K["fy"], K["cy"]], [0, 0, 1]], np.float32)
        F12 = (np.linalg.inv(Km).T @ tx @ R12 @ np.linalg.inv(Km)).astype(np.float32)
        Cw = -(R1w.T @ t1w)                                   # camera centre of KF1
        C2 = (R2w @ Cw + t2w).astype(np.float32)
        invz = np.float32(1.0) / C2[2]
        ex = np.float32(np.float32(np.float32(K["fx"]) * C2[0]) * invz + np.float32(K["cx"]))
        ey = np.float32(np.float32(np.float32(K["fy"]) * C2[1]) * invz + np.float32(K["cy"]))
        (k1, u1, d1, h1, f1), (k2, u2, d2, h2, f2) = views
        m12, nm = match_py.search_for_triangulation(k1, u1, d1, h1, f1, k2, u2, d2, h2, f2, F12, ex, ey, scale, sigma2,
                                                    only_stereo, True, 50)
        def csr(d):
            ids = sorted(d)
            offs = np.cumsum([0] + [len(d[k]) for k in ids]).astype(np.int32)
            return np.array(ids, np.uint32), offs, np.array([i for k in ids for i in d[k]], np.uint32)
        c1, c2 = csr(f1), csr(f2)
        np.savez_compressed(os.path.join(OUT, f"triang_pair{case}.npz"), kps1=k1, ur1=u1, desc1=d1, has_mp1=h1.astype(np.uint8),
                            kps2=k2, ur2=u2, desc2=d2, has_mp2=h2.astype(np.uint8), nodes1=c1[0], offs1=c1[1], idx1=c1[2],
                            nodes2=c2[0], offs2=c2[1], idx2=c2[2], F12=F12, ex=ex, ey=ey, scale=scale, sigma2=sigma2,
                            only_stereo=np.int32(only_stereo), matches12=m12, nmatches=np.int32(nm))
        print(f"triang_pair{case}: n {len(k1)}/{len(k2)} matches {nm}")


def make_fuse():
    """N1 fixture: the window search of ORBmatcher::Fuse — map points of a neighbour keyframe projected into a keyframe."""
    from oracle import orc
    from oracle.pyref import frame_py, match_py
    from psl_slam_b200._lib import FUSE_QUERY_DTYPE
    K = synth.ICL
    gray, depth, T = synth.sequence(6, 4)
    P = orb_cv2.OrbParams()
    scale = np.array(P.scale, np.float32)
    inv_sigma2 = (np.float32(1.0) / (scale * scale)).astype(np.float32)
    bounds = (np.float32(0), np.float32(0), np.float32(640), np.float32(480))
    rng = np.random.default_rng(41)
    for case, (a, b) in enumerate([(0, 3), (2, 1)]):
        ka, da = orc.orb_extract(gray[a])
        kb, db = orc.orb_extract(gray[b])
        xy = np.stack([ka["x"], ka["y"]], 1)
        _, dep = frame_py.stereo_from_rgbd(xy, frame_py.depth_to_float(depth[a]), K["bf"])
        world = frame_py.unproject(xy, dep, K, np.linalg.inv(T[a]))     # map points = keypoints of keyframe a
        xyb = np.stack([kb["x"], kb["y"]], 1)
        urb, _ = frame_py.stereo_from_rgbd(xyb, frame_py.depth_to_float(depth[b]), K["bf"])
        urb = np.where(rng.random(len(urb)) < 0.6, urb, -1.0).astype(np.float32)
        Tb = T[b].astype(np.float32)
        q = np.zeros(len(ka), FUSE_QUERY_DTYPE)
        for i in range(len(ka)):
            if not np.isfinite(world[i]).all() or dep[i] <= 0:
                continue
            pc = (Tb[:3, :3].astype(np.float64) @ world[i] + Tb[:3, 3]).astype(np.float32)
            if pc[2] < 0:
                continue
            invz = np.float32(1) / pc[2]
            u = np.float32(np.float32(K["fx"]) * np.float32(pc[0] * invz) + np.float32(K["cx"]))
            v = np.float32(np.float32(K["fy"]) * np.float32(pc[1] * invz) + np.float32(K["cy"]))
            if not (0 <= u < 640 and 0 <= v < 480):
                continue
            lvl = int(np.clip(ka["octave"][i] + rng.integers(-1, 2), 0, 7))
            q[i] = (u, v, np.float32(u - np.float32(K["bf"]) * invz), np.float32(3.0) * scale[lvl], lvl, 1)
        kun = np.stack([kb["x"], kb["y"], kb["size"], kb["angle"], kb["response"]], 1).astype(np.float32)
        fv = match_py.FrameView(kun, kb["octave"], urb, db, *bounds)
        bi, bd = match_py.fuse_search(fv, q, da, inv_sigma2, 50)
        np.savez_compressed(os.path.join(OUT, f"fuse_pair{case}.npz"), kps=kb, desc=db, u_right=urb,
                            bounds=np.array(bounds, np.float32), queries=q, qdesc=da, inv_sigma2=inv_sigma2, best_idx=bi,
                            best_dist=bd)
        print(f"fuse_pair{case}: queries {int((q['flags'] & 1).sum())} fused {(bi >= 0).sum()}")


def make_loop():
    """N1 fixtures, second batch (loop closing + monocular initialisation): SearchByBoW(KF, KF), SearchBySim3, the Sim3
    forms of Fuse and SearchByProjection, SearchForInitialization — two keyframe pairs of the synthetic sequence."""
    from oracle import orc
    from oracle.pyref import frame_py, match_py
    from psl_slam_b200._lib import FUSE_QUERY_DTYPE
    K = synth.ICL
    gray, depth, T = synth.sequence(7, 4)
    P = orb_cv2.OrbParams()
    scale = np.array(P.scale, np.float32)
    bounds = (np.float32(0), np.float32(0), np.float32(640), np.float32(480))
    rng = np.random.default_rng(71)

    def view(i):
        kps, desc = orc.orb_extract(gray[i])
        xy = np.stack([kps["x"], kps["y"]], 1)
        _, dep = frame_py.stereo_from_rgbd(xy, frame_py.depth_to_float(depth[i]), K["bf"])
        world = frame_py.unproject(xy, dep, K, np.linalg.inv(T[i]))
        kun = np.stack([kps["x"], kps["y"], kps["size"], kps["angle"], kps["response"]], 1).astype(np.float32)
        return kps, desc, dep, world, match_py.FrameView(kun, kps["octave"], None, desc, *bounds)

    def project(world, dep, octave, Tcw, th, valid):
        """the projection half of the Sim3 matchers (Scw already divided by its scale), float32 as cv::Mat"""
        Tc = Tcw.astype(np.float32)
        q = np.zeros(len(world), FUSE_QUERY_DTYPE)
        for i in range(len(world)):
            if not valid[i] or not np.isfinite(world[i]).all() or dep[i] <= 0:
                continue
            pc = (Tc[:3, :3].astype(np.float64) @ world[i] + Tc[:3, 3]).astype(np.float32)
            if pc[2] < 0:
                continue
            invz = np.float32(1) / pc[2]
            u = np.float32(np.float32(K["fx"]) * np.float32(pc[0] * invz) + np.float32(K["cx"]))
            v = np.float32(np.float32(K["fy"]) * np.float32(pc[1] * invz) + np.float32(K["cy"]))
            if not (0 <= u < 640 and 0 <= v < 480):
                continue
            lvl = int(np.clip(octave[i] + rng.integers(-1, 2), 0, P.nlevels - 1))
            q[i] = (u, v, np.float32(u - np.float32(K["bf"]) * invz), np.float32(th) * scale[lvl], lvl, 1)
        return q

    def fvec(desc, drop):
        node = (desc[:, 0] & 0x3F).astype(np.uint32)
        d = {}
        for i, nd in enumerate(node):
            if nd % 7 != drop:
                d.setdefault(int(nd), []).append(i)
        return d

    def csr(d):
        ids = sorted(d)
        offs = np.cumsum([0] + [len(d[k]) for k in ids]).astype(np.int32)
        return np.array(ids, np.uint32), offs, np.array([i for k in ids for i in d[k]], np.uint32)

    for case, (a, b) in enumerate([(0, 1), (3, 2)]):
        ka, da, depa, wa, fva = view(a)
        kb, db, depb, wb, fvb = view(b)
        n1, n2 = len(ka), len(kb)
        # the Sim3 estimate = the true relative pose, slightly off
        Tb, Ta = T[b].copy(), T[a].copy()
        Tb[:3, 3] += rng.normal(0, 0.003, 3)
        Ta[:3, 3] += rng.normal(0, 0.003, 3)
        # --- SearchByBoW(KF1, KF2)
        v1, v2 = rng.random(n1) < 0.85, rng.random(n2) < 0.85
        f1, f2 = fvec(da, 3), fvec(db, 5)
        bow12, nbow = match_py.search_by_bow_kf(da, ka["angle"], v1, f1, db, kb["angle"], v2, f2, 0.75, 50, True)
        bow12n, nbown = match_py.search_by_bow_kf(da, ka["angle"], v1, f1, db, kb["angle"], v2, f2, 0.9, 80, False)
        c1, c2 = csr(f1), csr(f2)
        # --- SearchBySim3 (th = 7.5, LoopClosing.cc): points of either keyframe not matched by the BoW pass
        q12 = project(wa, depa, ka["octave"], Tb, 7.5, v1 & (bow12 < 0))
        taken2 = np.zeros(n2, bool)
        taken2[bow12[bow12 >= 0]] = True
        q21 = project(wb, depb, kb["octave"], Ta, 7.5, v2 & ~taken2)
        sim12, nsim = match_py.search_by_sim3(fva, fvb, q12, da, q21, db, 100)
        # --- Fuse(pKF, Scw, vpPoints, th = 4) and SearchByProjection(pKF, Scw, vpPoints, vpMatched, th = 10)
        qf = project(wa, depa, ka["octave"], Tb, 4.0, rng.random(n1) < 0.9)
        fbi, fbd = match_py.fuse_search_sim3(fvb, qf, da, 50)
        qp = project(wa, depa, ka["octave"], Tb, 10.0, rng.random(n1) < 0.9)
        held = rng.random(n2) < 0.3
        pas, npm = match_py.search_by_projection_sim3(fvb, qp, da, held, 50)
        # --- SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, 100)
        pm = np.stack([ka["x"], ka["y"]], 1).astype(np.float32)
        if case == 1:   # a later attempt: some entries already moved to an earlier frame's matches
            mv = rng.random(n1) < 0.4
            pm[mv] += rng.normal(0, 6, (int(mv.sum()), 2)).astype(np.float32)
        ini12, nini, pm_out = match_py.search_for_initialization(ka, da, pm, fvb, 100, 0.9, 50, True)
        ini12n, ninin, pm_outn = match_py.search_for_initialization(ka, da, pm, fvb, 40, 0.8, 70, False)
        np.savez_compressed(os.path.join(OUT, f"loop_pair{case}.npz"), kps1=ka, desc1=da, kps2=kb, desc2=db,
                            bounds=np.array(bounds, np.float32), valid1=v1.astype(np.uint8), valid2=v2.astype(np.uint8),
                            nodes1=c1[0], offs1=c1[1], idx1=c1[2], nodes2=c2[0], offs2=c2[1], idx2=c2[2],
                            bow12=bow12, nbow=np.int32(nbow), bow12_noori=bow12n, nbow_noori=np.int32(nbown),
                            q12=q12, q21=q21, sim12=sim12, nsim=np.int32(nsim),
                            qfuse=qf, fuse_idx=fbi, fuse_dist=fbd,
                            qproj=qp, held=held.astype(np.uint8), proj_assign=pas, nproj=np.int32(npm),
                            prev_matched=pm, ini12=ini12, nini=np.int32(nini), prev_out=pm_out,
                            ini12_b=ini12n, nini_b=np.int32(ninin), prev_out_b=pm_outn)
        print(f"loop_pair{case}: n {n1}/{n2} bow {nbow}/{nbown} sim3 {nsim} (of {int((q12['flags'] & 1).sum())}/"
              f"{int((q21['flags'] & 1).sum())}) fuse {(fbi >= 0).sum()} proj {npm} init {nini}/{ninin}")


def make_line():
    """cfg-3 shaped fixtures: LSD (real cv2) -> merge -> top-N -> LBD (real cv2 blur/Sobel) -> line equations."""
    from oracle.pyref import line_py
    from psl_slam_b200._lib import KEYLINE_DTYPE
    cases = [("line_lowtex3", synth.make_lowtex(3), 200), ("line_lowtex5", synth.make_lowtex(5), 200),
             ("line_lowtex7_320x240", synth.make_lowtex(7, 320, 240), 200),
             ("line_textured_top60", synth.sequence(1, 1)[0][0], 60),
             ("line_flat", np.full((240, 320), 90, np.uint8), 200)]
    for name, img, nfeat in cases:
        raw, merged, kls, des, bits, eq = line_py.line_extract(img, nfeat)
        kl = np.zeros(len(kls), KEYLINE_DTYPE)
        for i, k in enumerate(kls):
            for f in KEYLINE_DTYPE.names:
                kl[i][f] = k[f]
        np.savez_compressed(os.path.join(OUT, name + ".npz"), image=img, nfeatures=np.int32(nfeat), lsd_raw=raw,
                            merged=merged, keylines=kl, lbd=des, ldesc=bits, lineeq=eq)
        print(name, img.shape, "raw", len(raw), "merged", len(merged), "kept", len(kl))


def make_linematch():
    """Line matcher fixtures: real extracted lines of frame pairs (oracle extraction, pinned above) and the
    independent Python restatement of the matchers over cv2.BFMatcher."""
    from oracle import orc
    from oracle.pyref import linematch_py as lm
    from psl_slam_b200._lib import LINE_QUERY_DTYPE
    g, _, _ = synth.sequence(2, 2)
    low = synth.make_lowtex(21)
    low2 = np.roll(np.roll(low, 3, axis=1), -2, axis=0)
    low2 = np.clip(low2.astype(np.int32) + np.rint(np.random.default_rng(5).normal(0, 1.0, low.shape)).astype(np.int32),
                   0, 255).astype(np.uint8)
    for pair, (a, b) in enumerate([(g[0], g[1]), (low, low2)]):
        rng = np.random.default_rng(100 + pair)
        kl_l, d_l, eq_l, _ = orc.line_extract(a, 200)
        kl_c, d_c, eq_c, _ = orc.line_extract(b, 200)
        bounds = np.array([0, 0, 640, 480], np.float32)
        has_ml = (rng.random(len(kl_l)) < 0.85).astype(np.uint8)
        nnr12, nnr_n = lm.match_nnr(d_l, d_c, 0.95)
        geom, geom_n = lm.search_geom(kl_l, d_l, has_ml, kl_c, d_c, bounds, 0.95)
        bf = lm.frame_bf_match(d_l, d_c, 0.95, 50)
        dbl, dbl_n = lm.search_double(d_l, d_c, 0.95, 50)
        # projection queries: last-frame lines "projected" with a small displacement
        nq = len(kl_l)
        q = np.zeros(nq, LINE_QUERY_DTYPE)
        jit = rng.normal(0, 1.5, (nq, 4)).astype(np.float32)
        q["x1"], q["y1"] = kl_l["start_x"] + jit[:, 0], kl_l["start_y"] + jit[:, 1]
        q["x2"], q["y2"] = kl_l["end_x"] + jit[:, 2], kl_l["end_y"] + jit[:, 3]
        q["radius"] = 6.0
        q["sx"], q["sy"], q["ex"], q["ey"] = kl_l["s_oct_x"], kl_l["s_oct_y"], kl_l["e_oct_x"], kl_l["e_oct_y"]
        q["length"] = kl_l["line_length"]
        q["flags"] = (has_ml.astype(np.uint32)) | (2 * (rng.random(nq) < 0.8)).astype(np.uint32)
        # synthetic 3-D lines of the current frame (z = 2 plane + noise) and map-line normals
        l3d = np.zeros((len(kl_c), 6))
        for i, k in enumerate(kl_c):
            l3d[i] = [(k["start_x"] - 320) / 240.0, (k["start_y"] - 240) / 240.0, 2.0 + rng.normal(0, .02),
                      (k["end_x"] - 320) / 240.0, (k["end_y"] - 240) / 240.0, 2.0 + rng.normal(0, .02)]
        for i, k in enumerate(kl_l):
            v = np.array([k["start_x"] - k["end_x"], k["start_y"] - k["end_y"], 0.0]) / 240.0
            q["normal"][i] = v + rng.normal(0, 0.02, 3) * np.linalg.norm(v)
        fr = lm.LineFrame(kl_c, d_c, eq_c, l3d, bounds)
        claimed = (rng.random(len(kl_c)) < 0.2).astype(np.uint8)
        a0, n0 = lm.search_by_projection(fr, q, d_l, claimed, 0, 0.95)
        q1 = q.copy()
        q1["radius"] = np.where(rng.random(nq) < 0.5, 5.0, 8.0).astype(np.float32) * np.float32(3.0)
        a1, n1 = lm.search_by_projection(fr, q1, d_l, claimed, 1, 0.8)
        area = [np.array(fr.in_area(q["x1"][i], q["y1"][i], q["x2"][i], q["y2"][i], 6.0, 0.96), np.int32)
                for i in range(min(nq, 12))]
        np.savez_compressed(os.path.join(OUT, f"linematch_pair{pair}.npz"), kl_last=kl_l, desc_last=d_l, kl_cur=kl_c,
                            desc_cur=d_c, eq_cur=eq_c, lines3d_cur=l3d, bounds=bounds, has_ml=has_ml, nnr12=nnr12,
                            nnr_n=np.int32(nnr_n), geom=geom, geom_n=np.int32(geom_n), bf=bf, dbl=dbl,
                            dbl_n=np.int32(dbl_n), queries0=q, queries1=q1, claimed=claimed, proj0=a0,
                            proj0_n=np.int32(n0), proj1=a1, proj1_n=np.int32(n1),
                            area_cat=np.concatenate(area) if area else np.zeros(0, np.int32),
                            area_len=np.array([len(x) for x in area], np.int32))
        print(f"linematch_pair{pair}: lines {len(kl_l)}/{len(kl_c)} nnr {nnr_n} geom {geom_n} bf {(bf >= 0).sum()} "
              f"double {dbl_n} proj0 {n0} proj1 {n1}")
    # structural-line / plane association
    rng = np.random.default_rng(77)
    n_ljl, n_map = 14, 40
    Tcw = np.eye(4, dtype=np.float32)
    Tcw[:3, :3] = synth.trajectory(3, 9)[2][:3, :3].astype(np.float32)
    Tcw[:3, 3] = [0.1, -0.05, 0.3]
    nrm = rng.normal(0, 1, (n_map, 3))
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    map_planes = np.concatenate([nrm, rng.uniform(-3, 3, (n_map, 1))], 1).astype(np.float32)
    planes_cam = np.zeros((n_ljl, 4), np.float32)
    pts = np.zeros((n_ljl, 15))
    Twc = np.linalg.inv(Tcw.astype(np.float64))
    for i in range(n_ljl):
        m = int(rng.integers(0, n_map))
        pw = map_planes[m].astype(np.float64) * (1 if i % 3 else -1)
        pw[3] += rng.normal(0, 0.03)
        planes_cam[i] = (Twc.T @ pw).astype(np.float32)  # pi_c = Twc^T pi_w
        # 5 world points near the plane
        for k in range(5):
            p = rng.normal(0, 1, 3)
            p -= (pw[:3] @ p + pw[3]) * pw[:3]
            pts[i, 3 * k:3 * k + 3] = p + rng.normal(0, 0.02, 3)
    bad = (rng.random(n_map) < 0.15).astype(np.uint8)
    a_0, n_0 = lm.plane_assoc(planes_cam, pts, Tcw, map_planes, bad, 0.1, 0.86, 0)
    a_1, n_1 = lm.plane_assoc(planes_cam, pts, Tcw, map_planes, None, 0.1, 0.86, 1)
    np.savez_compressed(os.path.join(OUT, "plane_assoc.npz"), planes_cam=planes_cam, pts=pts, Tcw=Tcw,
                        map_planes=map_planes, map_bad=bad, assign0=a_0, n0=np.int32(n_0), assign1=a_1, n1=np.int32(n_1))
    print("plane_assoc: mode0", n_0, a_0.tolist(), "mode1", n_1, a_1.tolist())


def make_linetriang_new():
    """N1 fixture: LSDmatcher::SearchForTriangulationNew (epipolar-overlap variant) between two keyframes of the
    synthetic sequence; F21 / F12 as LSDmatcher::ComputeF12 does (float32 cv::Mat products)."""
    from oracle import orc
    from oracle.pyref import linematch_py as lm
    K = synth.ICL
    Km = np.array([[K["fx"], 0, K["cx"]], [0, K["fy"], K["cy"]], [0, 0, 1]], np.float32)

    def f12(T1, T2):   # LSDmatcher.cpp:826-845
        T1, T2 = T1.astype(np.float32), T2.astype(np.float32)
        R1w, t1w, R2w, t2w = T1[:3, :3], T1[:3, 3], T2[:3, :3], T2[:3, 3]
        R12 = (R1w @ R2w.T).astype(np.float32)
        t12 = (-R1w @ R2w.T @ t2w + t1w).astype(np.float32)
        tx = np.array([[0, -t12[2], t12[1]], [t12[2], 0, -t12[0]], [-t12[1], t12[0], 0]], np.float32)
        return (np.linalg.inv(Km).T @ tx @ R12 @ np.linalg.inv(Km)).astype(np.float32)

    gray, _, T = synth.sequence(3, 6)
    for case, (a, b, dbl, nnr) in enumerate([(0, 1, True, 0.95), (1, 4, False, 0.9)]):
        rng = np.random.default_rng(300 + case)
        k1, d1, e1, _ = orc.line_extract(gray[a], 200)
        k2, d2, e2, _ = orc.line_extract(gray[b], 200)
        ml1, ml2 = (rng.random(len(k1)) < 0.2).astype(np.uint8), (rng.random(len(k2)) < 0.2).astype(np.uint8)
        F21, F12 = f12(T[b], T[a]), f12(T[a], T[b])
        m, n = lm.search_for_triangulation_new(d1, k1, e1, ml1, d2, k2, e2, ml2, F21, F12, nnr, 50, dbl)
        m2, n2 = lm.search_for_triangulation_new(d1, k1, e1, ml1, d2, k2, e2, ml2, F21, F12, 0.99, 90, not dbl)
        np.savez_compressed(os.path.join(OUT, f"linetriangnew_pair{case}.npz"), kl1=k1, desc1=d1, func1=e1, ml1=ml1, kl2=k2,
                            desc2=d2, func2=e2, ml2=ml2, F21=F21, F12=F12, nn_ratio=np.float32(nnr), is_double=np.int32(dbl),
                            pairs=m, npairs=np.int32(n), pairs_b=m2, npairs_b=np.int32(n2))
        print(f"linetriangnew_pair{case}: lines {len(k1)}/{len(k2)} pairs {n} / {n2}")


def make_linefuse():
    """LSDmatcher::Fuse window search: the lines of the linematch pairs as KeyFrame lines, the last frame's lines as
    projected MapLines; kf_desc plays pKF->mDescriptors (rows near the line descriptors so that some fuse)."""
    from oracle.pyref import linematch_py as lm
    from psl_slam_b200._lib import LINE_FUSE_QUERY_DTYPE
    for pair in (0, 1):
        g = np.load(os.path.join(OUT, f"linematch_pair{pair}.npz"))
        rng = np.random.default_rng(300 + pair)
        kl = g["kl_cur"].copy()
        kl["octave"] = rng.integers(0, 3, len(kl))          # exercise the level gate
        n = len(kl)
        kf_desc = rng.integers(0, 256, (max(n + 40, 64), 32), dtype=np.uint8)
        flips = rng.integers(0, 256, (n, 32), dtype=np.uint8) & rng.integers(0, 256, (n, 32), dtype=np.uint8) & \
            rng.integers(0, 256, (n, 32), dtype=np.uint8)   # ~1/8 of the bits
        kf_desc[:n] = g["desc_cur"] ^ np.where(rng.random((n, 1)) < 0.7, flips & rng.integers(0, 256, (n, 32), dtype=np.uint8), flips)
        kl_l, d_l = g["kl_last"], g["desc_last"]
        nq = len(kl_l)
        q = np.zeros(nq, LINE_FUSE_QUERY_DTYPE)
        jit = rng.normal(0, 2.0, (nq, 4)).astype(np.float32)
        q["u1"], q["v1"] = kl_l["start_x"] + jit[:, 0], kl_l["start_y"] + jit[:, 1]
        q["u2"], q["v2"] = kl_l["end_x"] + jit[:, 2], kl_l["end_y"] + jit[:, 3]
        q["radius"] = (3.0 * np.float32(1.2) ** rng.integers(0, 3, nq)).astype(np.float32) * np.float32(4.0)
        q["pred_level"] = rng.integers(-1, 4, nq)
        q["flags"] = (rng.random(nq) < 0.9).astype(np.uint32)
        if nq > 2:  # degenerate projection (both endpoints equal): NaN direction passes the cosine gate
            q["u2"][1], q["v2"][1] = q["u1"][1], q["v1"][1]
        qd = d_l.copy()
        # some MapLine descriptors equal to the row the reference compares against
        for i in range(0, min(nq, n), 5):
            qd[i] = kf_desc[i % n]
        bi, bd = lm.line_fuse(kl, kf_desc, q, qd, 0.998, 50)
        bi2, bd2 = lm.line_fuse(kl, kf_desc, q, qd, 0.9, 80)
        area = [np.array(lm.lines_in_area(kl, q["u1"][i], q["v1"][i], q["u2"][i], q["v2"][i], q["radius"][i], 0.998), np.int32)
                for i in range(min(nq, 16))]
        np.savez_compressed(os.path.join(OUT, f"linefuse_pair{pair}.npz"), kl=kl, kf_desc=kf_desc, queries=q, qdesc=qd,
                            best_idx=bi, best_dist=bd, best_idx_loose=bi2, best_dist_loose=bd2,
                            area_cat=np.concatenate(area) if area else np.zeros(0, np.int32),
                            area_len=np.array([len(x) for x in area], np.int32))
        print(f"linefuse_pair{pair}: lines {n} queries {int((q['flags'] & 1).sum())} fused {(bi >= 0).sum()} "
              f"loose {(bi2 >= 0).sum()} with candidates {(bd < 256).sum()}")


def make_undistort():
    """Frame::UndistortKeyPoints / ComputeImageBounds: the real cv2.undistortPoints on the ORB keypoints of a golden
    frame and on random points (inside, outside, on the corners), TUM1 / TUM2-like 5- and 4-coefficient cameras."""
    import cv2
    g = np.load(os.path.join(OUT, "orb_vga_seed1.npz"))
    rng = np.random.default_rng(9)
    pts_kp = np.ascontiguousarray(g["kps"][:, :2], np.float32)   # x, y of the final keypoints
    pts_rnd = np.concatenate([rng.uniform(-80, 760, (4000, 2)), rng.uniform(0, 640, (2000, 2)) * [1, 0.75],
                              [[0, 0], [640, 0], [0, 480], [640, 480], [318.64304, 255.31399]]]).astype(np.float32)
    cams = {"tum1": (517.306408, 516.469215, 318.643040, 255.313989, 0.262383, -0.953104, -0.005358, 0.002628, 1.163314),
            "tum2": (520.908620, 521.007327, 325.141442, 249.701764, 0.231222, -0.784899, -0.003257, -0.000105, 0.917205),
            "k1only": (481.2, 480.0, 319.5, 239.5, -0.28, 0.0, 0.0, 0.0, 0.0),
            "strong": (300.0, 300.0, 320.0, 240.0, -0.45, 0.25, 0.01, -0.008, 0.0)}
    out = {"pts_kp": pts_kp, "pts_rnd": pts_rnd}
    for name, c in cams.items():
        K = np.array([[c[0], 0, c[2]], [0, c[1], c[3]], [0, 0, 1]], np.float32)
        D = np.array(c[4:], np.float32)
        out[f"cam_{name}"] = np.array(c, np.float32)
        out[f"kp_{name}"] = cv2.undistortPoints(pts_kp.reshape(-1, 1, 2), K, D, None, K).reshape(-1, 2)
        out[f"rnd_{name}"] = cv2.undistortPoints(pts_rnd.reshape(-1, 1, 2), K, D, None, K).reshape(-1, 2)
        corners = np.array([[0, 0], [640, 0], [0, 480], [640, 480]], np.float32)
        m = cv2.undistortPoints(corners.reshape(-1, 1, 2), K, D, None, K).reshape(-1, 2)
        out[f"bounds_{name}"] = np.array([min(m[0, 0], m[2, 0]), min(m[0, 1], m[1, 1]), max(m[1, 0], m[3, 0]),
                                          max(m[2, 1], m[3, 1])], np.float32)   # Frame.cc:1150-1153 as (minX, minY, maxX, maxY)
        print(f"undistort {name}: max shift {np.abs(out[f'rnd_{name}'] - pts_rnd).max():.2f} px, bounds {out[f'bounds_{name}']}")
    np.savez_compressed(os.path.join(OUT, "undistort.npz"), **out)


def make_planes():
    """Frame::ExtractLSD plane hypotheses: synthetic structural lines on a few 3-D planes in the camera frame (coplanar
    intersecting pairs, repeated planes for OldPlane, skew pairs, lines without 3-D, parallel directions -> NaN)."""
    from oracle.pyref import linematch_py as lm
    from psl_slam_b200._lib import JUNCTION_DTYPE, KEYLINE_DTYPE
    K = synth.ICL
    for case in (0, 1, 2):
        rng = np.random.default_rng(700 + min(case, 1))   # case 2 = case 1 with the degenerate pair first
        n_planes, per = (5, 6) if case == 0 else (9, 8)
        kl, eq, l3, js = [], [], [], []
        def project(P):
            return K["fx"] * P[0] / P[2] + K["cx"], K["fy"] * P[1] / P[2] + K["cy"]
        for pi in range(n_planes):
            n = rng.normal(0, 1, 3); n /= np.linalg.norm(n)
            c = np.array([rng.uniform(-1, 1), rng.uniform(-0.8, 0.8), rng.uniform(1.5, 4.0)])
            u = np.cross(n, [0, 0, 1.0]); u /= np.linalg.norm(u)
            v = np.cross(n, u)
            first = len(kl)
            for li in range(per):
                a = rng.uniform(0, np.pi)
                dirv = np.cos(a) * u + np.sin(a) * v
                mid = c + rng.uniform(-0.3, 0.3) * u + rng.uniform(-0.3, 0.3) * v
                noise = rng.normal(0, 0.004 if li % 3 else 0.03, (2, 3))    # some lines leave the plane by > 5 cm
                A, B = mid - 0.4 * dirv + noise[0], mid + 0.4 * dirv + noise[1]
                k = np.zeros((), KEYLINE_DTYPE)
                (k["start_x"], k["start_y"]), (k["end_x"], k["end_y"]) = project(A), project(B)
                kl.append(k)
                d = (B - A).astype(np.float32)
                eq.append(d / np.float32(np.sqrt(np.float32(d @ d))))
                l3.append(np.concatenate([A, B]))
            for a_ in range(first, first + per):
                for b_ in range(a_ + 1, first + per):
                    if rng.random() < 0.45:
                        A1, B1, A2 = l3[a_][:3], l3[a_][3:], l3[b_][:3]
                        X = A1 + rng.uniform(0, 1) * (B1 - A1) * 0.5 + 0.5 * (A2 - A1) * rng.uniform(0, 0.2)
                        X = X + rng.normal(0, 0.003, 3)
                        j = np.zeros((), JUNCTION_DTYPE)
                        j["l1"], j["l2"] = a_, b_
                        j["cross2d_x"], j["cross2d_y"] = project(X)
                        j["cross3d"] = X
                        js.append(j)
        kl, eq, l3, js = np.array(kl), np.array(eq, np.float32), np.array(l3), np.array(js)
        # lines without a 3-D fit (isLineGood left them at their initial values), a zeroed direction, a skew pair and a
        # parallel pair (normal = 0 / 0)
        eq[3] = (-1, -1, -1); l3[3] = 0
        eq[7] = 0
        extra = np.zeros(3, JUNCTION_DTYPE)
        extra[0]["l1"], extra[0]["l2"] = 0, per + 1                      # lines of two different planes
        extra[1]["l1"], extra[1]["l2"] = 1, 1                            # a line with itself: parallel
        extra[2]["l1"], extra[2]["l2"] = 2, 3
        for e in extra:
            e["cross3d"] = l3[e["l1"]][:3]
        js = np.concatenate([js[: len(js) // 2], extra, js[len(js) // 2:]])
        js = js[rng.permutation(len(js))]
        if case == 2:  # the degenerate pair first: with no plane kept yet its NaN plane is accepted (and then, failing
            # every comparison of OldPlane, makes each later hypothesis look old) -- the reference does the same
            par = int(np.flatnonzero((js["l1"] == 1) & (js["l2"] == 1))[0])
            js = np.concatenate([js[par:par + 1], js[:par], js[par + 1:]])
        le, pl, nr, ow = lm.plane_hypotheses(kl, eq, l3, js)
        np.savez_compressed(os.path.join(OUT, f"planes_case{case}.npz"), kl=kl, line_eq=eq, lines3d=l3, junctions=js,
                            le_l=le, planes=pl, normals=nr, junction_of=ow)
        print(f"planes_case{case}: lines {len(kl)} junctions {len(js)} planes {len(pl)} "
              f"(nan planes {int(np.isnan(pl).any(1).sum())})")


def make_junctions():
    """N2 fixture: the junction detection of Frame::ExtractLSD (CPartiallyRecoverConnectivity over the real cv2
    primitives, oracle/pyref/junction_py.py) on the lines of the lines3d goldens; the 3-D cross points of
    Frame::convertFansToKeyLines from numpy.linalg (tolerance check only: the reference solves with Eigen)."""
    from oracle.pyref import junction_py
    from psl_slam_b200._lib import JUNCTION_DTYPE
    for case, (src, radius, thr) in enumerate([("lines3d_clean", 20.0, np.pi / 4), ("lines3d_noisy", 20.0, np.pi / 4),
                                               ("lines3d_lowtex", 20.0, np.pi / 4), ("lines3d_holes", 35.0, np.pi / 6)]):
        g = np.load(os.path.join(OUT, src + ".npz"))
        kl, l3 = g["kl"], g["lines3d"]
        L = np.stack([kl["start_x"], kl["start_y"], kl["end_x"], kl["end_y"]], 1).astype(np.float32)
        fans, raw = junction_py.fans(L, radius, thr, 640, 480)
        js = []
        for f in fans:
            i1, i2 = int(f[2]), int(f[3])
            ok, X = junction_py.cross3d_numpy(l3[i1], l3[i2])
            if ok and np.linalg.norm(X) > np.finfo(np.float64).eps:
                j = np.zeros((), JUNCTION_DTYPE)
                j["l1"], j["l2"], j["cross2d_x"], j["cross2d_y"], j["cross3d"] = i1, i2, f[0], f[1], X
                js.append(j)
        js = np.array(js, JUNCTION_DTYPE)
        np.savez_compressed(os.path.join(OUT, f"junctions_case{case}.npz"), kl=kl, lines3d=l3, radius=np.float32(radius),
                            fan_thr=np.float32(thr), size=np.array([640, 480], np.int32), fans=fans, fans_raw=raw,
                            junctions=js)
        print(f"junctions_case{case}: lines {len(kl)} raw fans {len(raw)} fans {len(fans)} junctions {len(js)}")


def make_pose():
    """N4 fixture (oracle side only): Optimizer::PoseOptimization with point edges — synthetic map points seen from a known
    pose with pixel noise, gross outliers and a perturbed pose prior; outputs of the independent numpy restatement
    (oracle/pyref/pose_py.py)."""
    from oracle import orc
    from oracle.pyref import pose_py
    K = synth.ICL
    fx, fy, cx, cy, bf = K["fx"], K["fy"], K["cx"], K["cy"], K["bf"]

    def rot(w):
        th = np.linalg.norm(w)
        k = w / th
        Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
        return np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx

    cases = [("pose_case0", 0, 600, 0.7, 0.10, 0.01, 0.03), ("pose_case1_mono", 1, 300, 0.0, 0.05, 0.02, 0.05),
             ("pose_case2_far_prior", 2, 800, 0.9, 0.20, 0.06, 0.25), ("pose_case3_few", 3, 8, 0.5, 0.0, 0.01, 0.02),
             ("pose_case4_two", 4, 2, 1.0, 0.0, 0.01, 0.02)]
    for name, seed, n, p_stereo, p_out, rot_err, t_err in cases:
        rng = np.random.default_rng(700 + seed)
        Rt, tt = rot(rng.normal(0, 0.05, 3)), rng.normal(0, 0.1, 3)
        Xc = np.stack([rng.uniform(-1.5, 1.5, n), rng.uniform(-1, 1, n), rng.uniform(0.8, 6, n)], 1)
        Xw = (Rt.T @ (Xc - tt).T).T
        u = fx * Xc[:, 0] / Xc[:, 2] + cx + rng.normal(0, 0.6, n)
        v = fy * Xc[:, 1] / Xc[:, 2] + cy + rng.normal(0, 0.6, n)
        ur = u - bf / Xc[:, 2] + rng.normal(0, 0.6, n)
        p = np.zeros(n + 5, orc.POSE_POINT_DTYPE)
        p["u"][:n], p["v"][:n] = u, v
        p["u_right"][:n] = np.where(rng.random(n) < p_stereo, ur, -1)
        p["inv_sigma2"][:n] = 1 / 1.2 ** (2 * rng.integers(0, 8, n))
        p["xw"][:n], p["yw"][:n], p["zw"][:n] = Xw[:, 0], Xw[:, 1], Xw[:, 2]
        p["flags"][:n] = 1                                     # the last five keypoints have no MapPoint
        bad = rng.random(n) < p_out
        p["u"][:n][bad] += rng.normal(0, 25, int(bad.sum())).astype(np.float32)
        p["v"][:n][bad] += rng.normal(0, 25, int(bad.sum())).astype(np.float32)
        p = p[rng.permutation(len(p))]
        T0 = np.eye(4, dtype=np.float32)
        T0[:3, :3] = rot(rng.normal(0, rot_err, 3)) @ Rt
        T0[:3, 3] = tt + rng.normal(0, t_err, 3)
        T, outl, cnt = pose_py.pose_optimization(T0, p, fx, fy, cx, cy, bf)
        Ttrue = np.eye(4)
        Ttrue[:3, :3], Ttrue[:3, 3] = Rt, tt
        np.savez_compressed(os.path.join(OUT, name + ".npz"), Tcw0=T0, pts=p, cam=np.array([fx, fy, cx, cy, bf], np.float32),
                            Tcw=T, outlier=outl, count=np.int32(cnt), Ttrue=Ttrue)
        print(f"{name}: points {int((p['flags'] & 1).sum())} inliers {cnt} outliers {int(outl.sum())} "
              f"|t - t_true| {np.abs(T[:3, 3] - tt).max():.2e} (prior {np.abs(T0[:3, 3] - tt).max():.2e})")


def make_pose_lil():
    """N4 with the structural-line edges: the point cases plus synthetic InsectLines — pairs of coplanar 3-D segments that meet
    in a cross point, seen from the true pose; the observed 2-D line equations go through the noisy projections of the end
    points (normalised as LINEextractor does, LineExtractor.cpp:352-363); some structural lines are gross outliers (wrong
    association) and some records carry no map InsectLine.  Outputs of oracle/pyref/pose_py.py."""
    from oracle import orc
    from oracle.pyref import pose_py
    K = synth.ICL
    fx, fy, cx, cy, bf = K["fx"], K["fy"], K["cx"], K["cy"], K["bf"]

    def rot(w):
        th = np.linalg.norm(w)
        k = w / th
        Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
        return np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx

    def proj(X):
        return np.stack([fx * X[..., 0] / X[..., 2] + cx, fy * X[..., 1] / X[..., 2] + cy], -1)

    # (name, seed, points, structural lines, share of outlier structural lines, prior rotation / translation error)
    cases = [("pose_lil_case0", 0, 400, 12, 0.2, 0.01, 0.03), ("pose_lil_case1_lines_only", 1, 0, 9, 0.0, 0.01, 0.02),
             ("pose_lil_case2_few_points", 2, 6, 20, 0.15, 0.02, 0.05), ("pose_lil_case3_far_prior", 3, 250, 16, 0.25, 0.05, 0.2)]
    for name, seed, n, m, p_out, rot_err, t_err in cases:
        rng = np.random.default_rng(900 + seed)
        Rt, tt = rot(rng.normal(0, 0.05, 3)), rng.normal(0, 0.1, 3)
        Xc = np.stack([rng.uniform(-1.5, 1.5, n), rng.uniform(-1, 1, n), rng.uniform(0.8, 6, n)], 1)
        Xw = (Rt.T @ (Xc - tt).T).T
        p = np.zeros(n + 3, orc.POSE_POINT_DTYPE)
        if n:
            uv = proj(Xc) + rng.normal(0, 0.6, (n, 2))
            p["u"][:n], p["v"][:n] = uv[:, 0], uv[:, 1]
            p["u_right"][:n] = np.where(rng.random(n) < 0.7, uv[:, 0] - bf / Xc[:, 2] + rng.normal(0, 0.6, n), -1)
            p["inv_sigma2"][:n] = 1 / 1.2 ** (2 * rng.integers(0, 8, n))
            p["xw"][:n], p["yw"][:n], p["zw"][:n] = Xw[:, 0], Xw[:, 1], Xw[:, 2]
            p["flags"][:n] = 1
            bad = rng.random(n) < 0.08
            p["u"][:n][bad] += rng.normal(0, 25, int(bad.sum())).astype(np.float32)
        # structural lines in camera coordinates: cross point, two directions in a random plane through it
        l = np.zeros(m + 2, orc.POSE_LIL_DTYPE)
        for i in range(m):
            c = np.array([rng.uniform(-1.2, 1.2), rng.uniform(-0.8, 0.8), rng.uniform(1.5, 5)])
            d1, d2 = rng.normal(0, 1, 3), rng.normal(0, 1, 3)
            d1, d2 = d1 / np.linalg.norm(d1), d2 / np.linalg.norm(d2)
            P = np.stack([c + rng.uniform(-0.1, 0.05) * d1, c + rng.uniform(0.3, 0.9) * d1,
                          c + rng.uniform(-0.1, 0.05) * d2, c + rng.uniform(0.3, 0.9) * d2, c])
            Pw = (Rt.T @ (P - tt).T).T
            l["line1"][i], l["line2"][i], l["cross"][i] = Pw[:2].reshape(6), Pw[2:4].reshape(6), Pw[4]
            uv = proj(P) + rng.normal(0, 0.5, (5, 2))
            if rng.random() < p_out:
                uv += rng.normal(0, 30, 2)             # a wrong association: everything shifted
            for key, a, b in (("obs1", uv[0], uv[1]), ("obs2", uv[2], uv[3])):
                le = np.cross(np.append(a, 1.0), np.append(b, 1.0))
                l[key][i] = le / np.sqrt(le[0] ** 2 + le[1] ** 2)
            l["ins"][i] = uv[4]
            l["flags"][i] = 1
        l = l[rng.permutation(len(l))]
        p = p[rng.permutation(len(p))]
        T0 = np.eye(4, dtype=np.float32)
        T0[:3, :3] = rot(rng.normal(0, rot_err, 3)) @ Rt
        T0[:3, 3] = tt + rng.normal(0, t_err, 3)
        T, outl, cnt, loutl = pose_py.pose_optimization(T0, p, fx, fy, cx, cy, bf, l)
        Ttrue = np.eye(4)
        Ttrue[:3, :3], Ttrue[:3, 3] = Rt, tt
        np.savez_compressed(os.path.join(OUT, name + ".npz"), Tcw0=T0, pts=p, lils=l, cam=np.array([fx, fy, cx, cy, bf], np.float32),
                            Tcw=T, outlier=outl, lil_outlier=loutl, count=np.int32(cnt), Ttrue=Ttrue)
        print(f"{name}: points {int((p['flags'] & 1).sum())} lils {int((l['flags'] & 1).sum())} count {cnt} point outliers "
              f"{int(outl.sum())} lil outliers {int(loutl.sum())} |t - t_true| {np.abs(T[:3, 3] - tt).max():.2e} "
              f"(prior {np.abs(T0[:3, 3] - tt).max():.2e})")


def make_lines3d():
    """Frame::isLineGood: the lines of the linematch pairs over the sequence's depth (clean, noisy, noisy with holes);
    both SVDs are the real cv2.SVDecomp (oracle/pyref/line3d_py.py)."""
    from oracle.pyref import line3d_py
    K = synth.ICL
    _, depth, _ = synth.sequence(2, 2)
    low_kl = np.load(os.path.join(OUT, "linematch_pair1.npz"))["kl_cur"]
    tex_kl = np.load(os.path.join(OUT, "linematch_pair0.npz"))["kl_cur"]
    rng = np.random.default_rng(900)
    base = (depth[1].astype(np.float32) * np.float32(1.0 / K["depth_factor"])).astype(np.float32)
    cases = [("clean", tex_kl, 0.0, 0.0, 7), ("noisy", tex_kl, 0.01, 0.05, 8), ("holes", tex_kl, 0.03, 0.2, 9),
             ("lowtex", low_kl, 0.02, 0.1, 10)]
    for name, kl, noise, holes, seed in cases:
        dep = (base * (1 + rng.normal(0, noise, base.shape))).astype(np.float32)
        dep[rng.random(base.shape) < holes] = 0
        if name == "lowtex":  # a depth step across the image: lines that straddle it need the RANSAC
            dep[:, 320:] = (dep[:, 320:] * np.float32(1.35)).astype(np.float32)
        # keep the fixture small: only the depth within 2 px of a line is ever read, the rest is zeroed
        import cv2
        mask = np.zeros(dep.shape, np.uint8)
        for k in kl:
            cv2.line(mask, (int(round(float(k["start_x"]))), int(round(float(k["start_y"])))),
                     (int(round(float(k["end_x"]))), int(round(float(k["end_y"])))), 1, thickness=5)
        dep = (dep * mask).astype(np.float32)
        l3, eq = line3d_py.is_line_good(kl, dep, K["fx"], K["fy"], K["cx"], K["cy"], seed)
        np.savez_compressed(os.path.join(OUT, f"lines3d_{name}.npz"), kl=kl, depth=dep,
                            cam=np.array([K["fx"], K["fy"], K["cx"], K["cy"]], np.float32), seed=np.uint32(seed),
                            lines3d=l3, line_eq=eq)
        print(f"lines3d_{name}: lines {len(kl)} with a 3-D line {int((np.abs(l3).sum(1) > 0).sum())}")


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    os.makedirs(OUT, exist_ok=True)
    if what in ("orb", "all"):
        make_orb()
    if what in ("match", "all"):
        make_match()
    if what in ("triang", "all"):
        make_triangulation()
    if what in ("fuse", "all"):
        make_fuse()
    if what in ("loop", "all"):
        make_loop()

    if what in ("line", "all"):
        make_line()
    if what in ("linematch", "all"):
        make_linematch()
    if what in ("linefuse", "all"):
        make_linefuse()
    if what in ("linetriangnew", "all"):
        make_linetriang_new()
    if what in ("pose", "all"):
        make_pose()
    if what in ("pose_lil", "all"):
        make_pose_lil()
    if what in ("undistort", "all"):
        make_undistort()
    if what in ("planes", "all"):
        make_planes()
    if what in ("lines3d", "all"):
        make_lines3d()
    if what in ("junctions", "all"):   # reads the lines3d fixtures
        make_junctions()
