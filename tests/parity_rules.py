"""The per-pixel parity rule of the GPU tests (north_star: 'rendered images, using the reference's per-pixel RNG seeding, match
within a stated per-pixel relative tolerance on identical paths').

TOLERANCE.  A pixel is IN TOLERANCE when every channel satisfies |gpu - oracle| <= ATOL_REL * max(oracle image) + RTOL * |oracle|
with RTOL = 2e-3, ATOL_REL = 1e-5.  (sinf / cosf / expf differ by ulps between CUDA and glibc, and the gather uses ex2.approx.)

IDENTICAL PATHS.  A pixel may be out of tolerance only if a DISCRETE decision on a path that feeds it went the other way in the two
renderers — Russian roulette, Fresnel reflect / refract, which primitive an edge-grazing ray hits, occluded or not, inside or
outside the gather radius.  The tests do not assume this, they check it: every out-of-tolerance pixel must be EXPLAINED by one of
the measured divergences below, and the fractions are printed.

  state      the final XORWOW state of the pixel's RNG stream differs from the oracle's (the stream is shared by the pixel's camera
             path, the photon thread / light subpath of the same index and the pixel's light samples, OptixRenderer.cpp:368-373), or
             the hit-point flags / position differ;
  PPM indirect   a photon that exists in only one of the two photon maps (slot-by-slot comparison of the photon buffers as emitted),
             or differs beyond tolerance, lies within the gather radius of the pixel's hit point; or a photon lies within 1e-4 r of
             the radius sphere itself (accept / reject decided by ulps);
  PPM direct     the set of unoccluded shadow samples of the pixel differs (4-bit mask);
  BDPT       the number of contributions that passed the occlusion test differs for the pixel (camera-path connections + light
             -tracing splats that landed on it), or a light vertex of a diverged path (state) projects onto the pixel in either
             renderer (its splat moved or changed).

`Probe` adapts the product (debug buffers over the C ABI) and the oracle to one interface.
"""
import numpy as np
from scipy.spatial import cKDTree

import oracle_lib as ol

RTOL, ATOL_REL = 2e-3, 1e-5


def out_of_tolerance(a, b, rtol=RTOL, atol_rel=ATOL_REL):
    """Per-pixel bool: some channel of a differs from b beyond the stated tolerance.  a, b: (..., 3)."""
    a, b = a.reshape(-1, 3), b.reshape(-1, 3)
    atol = atol_rel * float(np.abs(b).max() + 1e-30)
    return (np.abs(a - b) > atol + rtol * np.abs(b)).any(axis=1)


class ProductProbe:
    def __init__(self, r, W, H, pl):
        self.r, self.W, self.H, self.pl = r, W, H, pl
        self.rsW, self.rsH = max(pl[0], W), max(pl[1], H)

    def hitpoints(self):
        return self.r.debug_buffer("hitpoints", ol.HITPOINT_DT, self.W * self.H)

    def random_states(self):
        return self.r.debug_buffer("random_states", np.uint32, 6 * self.rsW * self.rsH).reshape(-1, 6)

    def photons_as_emitted(self):
        return self.r.debug_buffer("photons", ol.PHOTON_DT, self.pl[0] * self.pl[1] * 4)

    def indirect(self):
        return self.r.debug_buffer("indirect", np.float32, self.W * self.H * 3).reshape(-1, 3)

    def direct(self):
        return self.r.debug_buffer("direct", np.float32, self.W * self.H * 3).reshape(-1, 3)

    def direct_sample_mask(self):
        return self.r.debug_buffer("direct_sample_mask", np.uint8, self.W * self.H)

    def connection_counts(self):
        return self.r.debug_buffer("connection_counts", np.uint32, self.W * self.H)

    def light_vertex_positions(self):
        return self.r.light_vertex_positions()

    def output(self):
        return self.r.getOutputBuffer().reshape(-1, 3).copy()


class OracleProbe:
    def __init__(self, o, W, H, pl):
        self.o, self.W, self.H, self.pl = o, W, H, pl
        self.rsW, self.rsH = max(pl[0], W), max(pl[1], H)

    def hitpoints(self):
        return self.o.hitpoints()

    def random_states(self):
        return self.o.random_states()

    def photons_as_emitted(self):
        return self.o.photons_as_emitted()

    def indirect(self):
        return self.o.indirect().reshape(-1, 3)

    def direct(self):
        return self.o.direct().reshape(-1, 3)

    def direct_sample_mask(self):
        return self.o.direct_sample_mask()

    def connection_counts(self):
        return self.o.connection_counts()

    def light_vertex_positions(self):
        return self.o.light_vertex_positions()

    def output(self):
        return self.o.output().reshape(-1, 3).copy()


def state_divergence(a, b):
    """(rsH, rsW) bool: final RNG state differs."""
    sa, sb = a.random_states(), b.random_states()
    assert sa.shape == sb.shape, (sa.shape, sb.shape)
    return (sa != sb).any(axis=1).reshape(a.rsH, a.rsW)


def pixel_state_divergence(a, b):
    return state_divergence(a, b)[:a.H, :a.W].reshape(-1)


def camera_divergence(a, b):
    """Per pixel: RNG state, hit flags or (for diffuse hits) hit position / normal differ."""
    hpa, hpb = a.hitpoints(), b.hitpoints()
    div = pixel_state_divergence(a, b) | (hpa["flags"] != hpb["flags"])
    diffuse = ((hpb["flags"] & ol.PRD_HIT_NON_SPECULAR) != 0) & ~div
    scale = float(np.abs(hpb["position"][diffuse]).max()) if diffuse.any() else 1.0
    moved = ~np.isclose(hpa["position"], hpb["position"], rtol=1e-4, atol=1e-4 * scale).all(axis=1)
    turned = ~np.isclose(hpa["normal"], hpb["normal"], rtol=0, atol=1e-3).all(axis=1)
    return div | (diffuse & (moved | turned)), hpa, hpb


def ppm_explained(a, b, radius):
    """Masks (per pixel) of the PPM rule after one iteration run on both renderers.  Returns dict(camera, indirect, direct)."""
    cam, hpa, hpb = camera_divergence(a, b)
    pa, pb = a.photons_as_emitted(), b.photons_as_emitted()
    va, vb = pa["power"].max(axis=1) > 0, pb["power"].max(axis=1) > 0
    scale = float(np.abs(pb["position"][vb]).max()) if vb.any() else 1.0
    close = np.isclose(pa["position"], pb["position"], rtol=1e-4, atol=1e-4 * scale).all(axis=1) & \
        np.isclose(pa["power"], pb["power"], rtol=1e-3, atol=1e-6 * float(pb["power"].max() + 1e-30)).all(axis=1) & \
        np.isclose(pa["rayDirection"], pb["rayDirection"], rtol=0, atol=1e-3).all(axis=1)
    same = (va == vb) & (~vb | close)
    suspects = np.concatenate([pa["position"][~same & va], pb["position"][~same & vb]])
    diffuse = (hpb["flags"] & ol.PRD_HIT_NON_SPECULAR) != 0
    near = np.zeros(a.W * a.H, bool)
    edge = np.zeros(a.W * a.H, bool)
    idx = np.nonzero(diffuse)[0]
    if idx.size:
        pos = hpb["position"][idx].astype(np.float64)
        if suspects.shape[0]:
            tree = cKDTree(suspects.astype(np.float64))
            near[idx] = tree.query_ball_point(pos, radius * 1.001 + 1e-4 * scale, return_length=True) > 0
        if vb.any():
            ptree = cKDTree(pb["position"][vb].astype(np.float64))
            outer = ptree.query_ball_point(pos, radius * (1 + 1e-4) + 2e-6 * scale, return_length=True)
            inner = ptree.query_ball_point(pos, max(radius * (1 - 1e-4) - 2e-6 * scale, 0.0), return_length=True)
            edge[idx] = outer != inner
    mask_diff = a.direct_sample_mask() != b.direct_sample_mask()
    return dict(camera=cam, indirect=cam | near | edge, direct=cam | (diffuse & mask_diff), photons_differ=int((~same).sum()),
                photons=int(vb.sum()), state=pixel_state_divergence(a, b), near=near, edge=edge, mask=mask_diff & diffuse)


def project_to_pixel(cam, W, H, points):
    """Pixel hit by the segment light vertex -> eye (connectCameraT1, vcm/vcm.h:65-161), or -1; float32 like the device code."""
    f = np.float32
    eye, look = np.array(list(cam.eye), f), np.array(list(cam.lookdir), f)
    cu, cv = np.array(list(cam.camera_u), f), np.array(list(cam.camera_v), f)
    ips = np.array(list(cam.imagePlaneSize), f)
    to_cam = eye - points.astype(f)
    front = (to_cam * -look).sum(axis=1) > 0
    dist = np.linalg.norm(to_cam, axis=1).astype(f)
    d = to_cam / np.maximum(dist, f(1e-30))[:, None]
    cos_cam = (-d * (look / np.linalg.norm(look))).sum(axis=1)
    ipd = np.linalg.norm(look) / np.where(cos_cam != 0, cos_cam, 1)
    c2p = (eye + (-d * ipd[:, None])) - (eye + look)
    pu = (c2p * (cu / np.linalg.norm(cu))).sum(axis=1)
    pv = (c2p * (cv / np.linalg.norm(cv))).sum(axis=1)
    inside = front & (2 * np.abs(pu) < ips[0]) & (2 * np.abs(pv) < ips[1])
    fx, fy = (pu + 0.5 * ips[0]) / ips[0] * W, (pv + 0.5 * ips[1]) / ips[1] * H
    return inside, fx, fy


def _mark_projected(mask, W, H, inside, fx, fy, border=2e-3):
    """Marks the pixel each point falls in, and the neighbour too when the point is within `border` px of a pixel edge."""
    fx, fy = fx[inside], fy[inside]
    for ox in (-border, border):
        for oy in (-border, border):
            x = np.clip(np.floor(fx + ox).astype(np.int64), 0, W - 1)
            y = np.clip(np.floor(fy + oy).astype(np.int64), 0, H - 1)
            mask[y * W + x] = True


def bdpt_explained(a, b, cam):
    """Per-pixel mask of the BDPT rule after one iteration run on both renderers (connection counts must be enabled)."""
    W, H = a.W, a.H
    state = pixel_state_divergence(a, b)
    count = a.connection_counts() != b.connection_counts()
    splat = np.zeros(W * H, bool)
    for probe in (a, b):
        path, pos = probe.light_vertex_positions()
        sel = state[path]
        _mark_projected(splat, W, H, *project_to_pixel(cam, W, H, pos[sel]))
    return dict(all=state | count | splat, state=state, count=count, splat=splat)


def check_explained(name, bad, explained, extra=""):
    """Assert bad ⊂ explained; returns the one-line report the tests print."""
    n = bad.size
    unexplained = bad & ~explained
    line = "%s: %d px, out of tolerance %d (%.3f %%), of those explained %d, UNEXPLAINED %d; explained set covers %.2f %% of the image%s" % (
        name, n, int(bad.sum()), 100.0 * bad.mean(), int((bad & explained).sum()), int(unexplained.sum()), 100.0 * explained.mean(), extra)
    print(line)
    assert not unexplained.any(), line + " first unexplained pixels: %s" % np.nonzero(unexplained)[0][:10].tolist()
    return line


def run_lockstep(name, r, o, det, odet, method, W, H, pl, radii, cam_pod, max_explained=0.5, per_iteration=None):
    """Renders len(radii) iterations on the product `r` and the oracle `o` in lock-step and applies the rule to every iteration's
    buffers (PPM: indirect and direct; BDPT / PT: the increment of the sum buffer) and to the final sum buffer (explained = union
    over the iterations).  The explained set must stay below `max_explained` of the image, or the rule would explain anything.
    Returns dict(bad=fraction out of tolerance, explained=fraction explained, lines=[report lines])."""
    a, b = ProductProbe(r, W, H, pl), OracleProbe(o, W, H, pl)
    if method == 1:
        r.set_connection_counts(True)
    union = np.zeros(W * H, bool)
    prev_a = prev_b = np.zeros((W * H, 3), np.float32)
    lines = []
    for i, radius in enumerate(radii):
        r.renderNextIteration(i, i, radius, True, det)
        o.render_next_iteration(i, i, radius, odet)
        if method == 2:
            e = ppm_explained(a, b, radius)
            lines.append(check_explained("%s it%d indirect" % (name, i), out_of_tolerance(a.indirect(), b.indirect()), e["indirect"],
                                         "; photons differing %d of %d, pixels: state %.4f, suspect photon in range %.4f, photon on the radius sphere %.4f"
                                         % (e["photons_differ"], e["photons"], e["state"].mean(), e["near"].mean(), e["edge"].mean())))
            lines.append(check_explained("%s it%d direct" % (name, i), out_of_tolerance(a.direct(), b.direct()), e["direct"],
                                         "; occlusion mask differs %.4f" % e["mask"].mean()))
            union |= e["indirect"] | e["direct"]
        else:
            if method == 1:
                e = bdpt_explained(a, b, cam_pod)
                ex, extra = e["all"], "; state %.4f, connection count %.4f, splat of a diverged path %.4f" % (e["state"].mean(), e["count"].mean(), e["splat"].mean())
            else:
                ex, extra = pixel_state_divergence(a, b), ""
            oa, ob = a.output(), b.output()
            lines.append(check_explained("%s it%d increment" % (name, i), out_of_tolerance(oa - prev_a, ob - prev_b, rtol=2 * RTOL), ex, extra))
            prev_a, prev_b = oa, ob
            union |= ex
        if per_iteration is not None:
            per_iteration(i, a, b)
    oa, ob = a.output(), b.output()
    bad = out_of_tolerance(oa, ob)
    lines.append(check_explained("%s sum buffer after %d iterations" % (name, len(radii)), bad, union))
    assert union.mean() <= max_explained, "%s: the explained set covers %.1f %% of the image" % (name, 100 * union.mean())
    rel = float(np.abs(oa - ob).sum() / (np.abs(ob).sum() + 1e-30))
    return dict(bad=float(bad.mean()), explained=float(union.mean()), rel_l1=rel, lines=lines)
