#!/usr/bin/env python
"""Diagnostics for the per-pixel parity rule (tests/parity_rules.py): runs one case of tests/test_parity_rule_gpu.py in lock-step
on the product and the oracle and prints, for every out-of-tolerance pixel the rule does not explain, the values a person needs
to name the discrete decision that differed (photons at the radius sphere / at the tangent plane, hit flags, path signature).
Needs a GPU.   python scripts/diag_parity.py sponza 0.08 ppm|bdpt|pt    |   cornell ppm   |   cornellsmall <config> ppm"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oppositerenderer_b200 as product  # noqa: E402
import oracle_lib as oracle  # noqa: E402
import parity_rules as pr  # noqa: E402
import test_parity_rule_gpu as T  # noqa: E402

M = {"ppm": T.PPM, "bdpt": T.BDPT, "pt": T.PT}


def ppm_pixel_report(p, a, b, radius, e):
    hpa, hpb = a.hitpoints(), b.hitpoints()
    print("  pixel %d  indirect gpu %s  oracle %s" % (p, a.indirect()[p], b.indirect()[p]))
    print("    flags %08x / %08x  pos %s / %s  normal %s / %s" % (hpa["flags"][p], hpb["flags"][p], hpa["position"][p], hpb["position"][p], hpa["normal"][p], hpb["normal"][p]))
    print("    explained: camera %s near %s edge %s state %s" % (e["camera"][p], e["near"][p], e["edge"][p], e["state"][p]))
    r2 = np.float32(radius) * np.float32(radius)
    for tag, probe, hp in (("gpu", a, hpa), ("oracle", b, hpb)):
        ph = probe.photons_as_emitted()
        v = ph["power"].max(axis=1) > 0
        pos = ph["position"][v]
        diff = hp["position"][p][None, :] - pos
        d2 = (diff[:, 0] * diff[:, 0] + diff[:, 1] * diff[:, 1]) + diff[:, 2] * diff[:, 2]
        dots = -(ph["rayDirection"][v] * hp["normal"][p][None, :]).sum(axis=1)
        inside = d2 <= r2
        k = np.argsort(np.abs(d2 - r2))[:3]
        near_r = inside | (np.abs(d2 - r2) < 1e-3 * r2)
        kd = np.nonzero(near_r)[0]
        kd = kd[np.argsort(np.abs(dots[kd]))[:3]] if kd.size else kd
        print("    %s: photons inside r %d, accepted %d; closest to the sphere (d2/r2 - 1): %s; smallest |dot(-dir, n)| among those inside: %s"
              % (tag, int(inside.sum()), int((inside & (dots >= 0)).sum()), (d2[k] / r2 - 1).tolist(), dots[kd].tolist()))


def main():
    args = sys.argv[1:]
    kind = args[0]
    if kind in ("sponza", "conference"):
        detail, method = float(args[1]), M[args[2]]
        W, H = 96, 54
        pl = (96, 96) if method == T.PPM else (16, 16)
        scene, cam, ocam, r, o, _ = T._synthetic(product, oracle, kind, detail, W, H, pl, 8)
        radii = product.ppm_radius_schedule(scene.getSceneInitialPPMRadiusEstimate(), 2)
    else:
        name = "Cornell" if kind == "cornell" else "CornellSmall"
        config = 0 if kind == "cornell" else int(args[1])
        method = M[args[-1]]
        size = int(os.environ.get("DIAG_SIZE", "64" if method == T.PPM else "48"))
        W = H = size
        pl = (size, size) if method == T.PPM else (16, 16)
        if os.environ.get("DIAG_PHOTONS"):
            pl = (int(os.environ["DIAG_PHOTONS"]),) * 2
        scene, cam, ocam, r, o, _ = T._builtin(product, oracle, name, config, W, H, pl, 10)
        r0 = scene.getSceneInitialPPMRadiusEstimate() if name == "Cornell" else 0.05
        radii = product.ppm_radius_schedule(r0, int(os.environ.get("DIAG_ITERS", "3"))) if method == T.PPM else [1.0, 1.0, 1.0]
    det = product.RenderServerRenderRequestDetails(cam, scene.getSceneName(), method, W, H, 2.0 / 3.0)
    odet = oracle.make_details(ocam, method, W, H)
    a, b = pr.ProductProbe(r, W, H, pl), pr.OracleProbe(o, W, H, pl)
    if method != T.PPM:
        r.set_connection_counts(True)
    prev_a = prev_b = np.zeros((W * H, 3), np.float32)
    for i, radius in enumerate(radii):
        r.renderNextIteration(i, i, radius, True, det)
        o.render_next_iteration(i, i, radius, odet)
        if method == T.PPM:
            e = pr.ppm_explained(a, b, radius)
            bad = pr.out_of_tolerance(a.indirect(), b.indirect())
            slack, st = pr.ppm_indirect_slack(a, b, radius, e["camera"], e["hpa"], e["hpb"])
            ok, ratio = pr.within_tolerance_plus_slack(a.indirect(), b.indirect(), slack, st["drift"])
            un = np.nonzero(~ok & ~e["camera"])[0]
            print("it%d indirect: bad %d, outside tolerance + slack and not camera-divergent %d" % (i, bad.sum(), un.size))
            for p in un[:6]:
                ppm_pixel_report(p, a, b, radius, e)
                # brute-force recomputation of the pixel from each renderer's photons AS EMITTED and its own hit point
                for tag, probe, hp in (("gpu", a, e["hpa"]), ("oracle", b, e["hpb"])):
                    ph = probe.photons_as_emitted()
                    v = ph["power"].max(axis=1) > 0
                    n = ph.shape[0]
                    c = pr._pair_contribution(hp, ph, np.full(n, p), np.arange(n), v, radius, float(pl[0] * pl[1])).sum(axis=0)
                    print("    %s: recomputed from the photons as emitted %s, buffer %s, attenuation %s" % (tag, c, probe.indirect()[p], hp["attenuation"][p]))
                print("    slack %s drift %.2e ratio %.2f" % (slack[p], st["drift"][p], ratio[p]))
                pa_, pb_ = a.photons_as_emitted(), b.photons_as_emitted()
                for tag, ph, hp in (("gpu", pa_, e["hpa"]), ("oracle", pb_, e["hpb"])):
                    d = np.linalg.norm(ph["position"].astype(np.float64) - hp["position"][p], axis=1)
                    for sl in np.nonzero((d < radius) & (ph["power"].max(axis=1) > 0))[0][:8]:
                        print("      in range of the %s hit point: slot %d  gpu %s | oracle %s" % (tag, sl, pa_[sl], pb_[sl]))
            badd = pr.out_of_tolerance(a.direct(), b.direct())
            und = np.nonzero(badd & ~e["direct"])[0]
            print("it%d direct: bad %d unexplained %d" % (i, badd.sum(), und.size))
            for p in und[:6]:
                print("  pixel %d direct gpu %s oracle %s mask %x / %x" % (p, a.direct()[p], b.direct()[p], a.direct_sample_mask()[p], b.direct_sample_mask()[p]))
        else:
            ex = pr.bdpt_explained(a, b, ocam)["all"] if method == T.BDPT else pr.path_tracing_explained(a, b)
            oa, ob = a.output(), b.output()
            bad = pr.out_of_tolerance(oa - prev_a, ob - prev_b, rtol=2 * pr.RTOL)
            un = np.nonzero(bad & ~ex)[0]
            print("it%d increment: bad %d unexplained %d" % (i, bad.sum(), un.size))
            hpa, hpb = a.hitpoints(), b.hitpoints()
            for p in un[:12]:
                print("  pixel %d gpu %s oracle %s  flags %08x / %08x att %s / %s" % (p, (oa - prev_a)[p], (ob - prev_b)[p], hpa["flags"][p], hpb["flags"][p],
                                                                                    hpa["attenuation"][p], hpb["attenuation"][p]))
                if method == T.PT:
                    print("      signature %s / %s  radiance %s / %s normal %s / %s pos %s / %s" % (a.path_signature()[p], b.path_signature()[p], hpa["radiance"][p], hpb["radiance"][p],
                                                                                  hpa["normal"][p], hpb["normal"][p], hpa["position"][p], hpb["position"][p]))
            prev_a, prev_b = oa, ob


if __name__ == "__main__":
    main()
