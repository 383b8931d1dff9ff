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
  PPM indirect   (quantitative, ppm_indirect_slack) the tolerance of a pixel is widened by the summed |C_gpu - C_oracle| over the photon
             slots that differ between the two maps and lie in range, C being the slot's contribution evaluated with each
             renderer's own photon and hit point — and the gather itself is checked against the oracle's gather on the product's
             own inputs with no exception (gather_stage_exact).  The mask form below is kept for the sum buffer:
             a photon that exists in only one of the two photon maps (slot-by-slot comparison of the photon buffers as emitted),
             or differs beyond tolerance (position by more than 2e-4 r — ulps of sinf / cosf grow along a multi-bounce walk, and
             the gather weight falls to zero at the radius — power by 1e-3, direction by 1e-3), lies within the gather radius of
             the pixel's hit point; or a photon lies within 1e-4 r of
             the radius sphere itself (accept / reject decided by ulps);
  PPM direct     the photon thread that shares the pixel's RNG stream reached another depth or stored another number of photons (the pixel's light
             samples come after it in the stream; the final state alone can hide this: an area-light sample takes three draws, a point-light
             sample one, so the direct pass can consume the opposite difference); the hit point lies in the plane of an area light (its samples arrive in the tangent plane: n.l / d^2 is 0/0-conditioned);
             the set of unoccluded shadow samples of the pixel differs (4-bit mask), or the hit's normal / attenuation differ by more than
             2e-5 / 2e-4 (textured and normal-mapped hits: 1/256 filter weights of the texture unit);
  PATH_TRACING   the path signature differs: per path segment, the hit flags, whether the light sample contributed (occlusion) and
             whether Russian roulette ended the path (the reference draws these from a stream whose final state it discards); or the
             last vertex of the path moved by more than 1e-5 of the scene size (same decisions, drifted geometry);
  BDPT       the light or camera subpath of the pixel's index met a different sequence of material kinds (path signature), or stored its
             light vertices more than 1e-5 of the scene size away (same decisions, drifted geometry); the number of contributions that passed the occlusion test differs for the pixel (camera-path connections + light
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

    def photon_path_lengths(self):
        return self.r.debug_buffer("photon_path_lengths", np.uint32, self.pl[0] * self.pl[1])

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

    def path_signature(self):
        return self.r.debug_buffer("pt_signature", np.uint32, self.W * self.H)

    def path_signatures(self):
        a = self.r.debug_buffer("path_signatures", np.uint32, 2 * self.W * self.H)
        return a[:self.W * self.H], a[self.W * self.H:]

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

    def photon_path_lengths(self):
        return self.o.photon_path_lengths()

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

    def path_signature(self):
        return self.o.pt_signature()

    def path_signatures(self):
        return self.o.path_signatures()

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


def photon_thread_divergence(a, b, pa=None, pb=None):
    """Per PIXEL: the photon thread that shares the pixel's RNG stream (same launch index, OptixRenderer.cpp:368-373) walked differently in
    the two renderers — another depth, or another number of stored photons.  Such a thread consumed a different number of draws, so the
    pixel's light samples, which come next in the stream, use shifted random numbers; when the direct pass happens to consume the
    opposite difference (an area light takes three draws, a point light one) the FINAL state is equal again and alone would not show it."""
    pw, ph = a.pl
    la, lb = a.photon_path_lengths(), b.photon_path_lengths()
    pa = a.photons_as_emitted() if pa is None else pa
    pb = b.photons_as_emitted() if pb is None else pb
    na = (pa["power"].max(axis=1) > 0).reshape(-1, 4).sum(axis=1)
    nb = (pb["power"].max(axis=1) > 0).reshape(-1, 4).sum(axis=1)
    div = ((la != lb) | (na != nb)).reshape(ph, pw)
    out = np.zeros((a.H, a.W), bool)
    hh, ww = min(a.H, ph), min(a.W, pw)
    out[:hh, :ww] = div[:hh, :ww]
    return out.reshape(-1)


def shading_inputs_differ(hpa, hpb):
    """Per pixel: the hit point's normal or attenuation differs measurably between the renderers although the camera path made the same
    decisions — normal-mapped / textured hits: the texture unit filters with 1/256 weights, so an ulp in (u, v) can move a weight by one
    step.  The direct term scales with n.l and the attenuation, so a 1e-4 change of the normal at grazing light exceeds 2e-3."""
    dn = np.abs(hpa["normal"].astype(np.float64) - hpb["normal"]).max(axis=1)
    att = np.maximum(np.abs(hpb["attenuation"]).max(axis=1).astype(np.float64), 1e-30)
    da = np.abs(hpa["attenuation"].astype(np.float64) - hpb["attenuation"]).max(axis=1) / att
    return (dn > 2e-5) | (da > 2e-4)


def coplanar_with_area_light(desc, hp):
    """Per pixel: the hit point lies in the plane of an area light (within 1e-4 of the scene size) — a ceiling next to a ceiling lamp.
    Its light samples then arrive in the tangent plane: n.l -> 0 while 1 / d^2 grows, the sample's factor n.l / (pi d^2) is 0/0-conditioned
    and an ulp of the hit point changes it by tens of per cent (DirectRadianceEstimation.cu:54-70, light.h:29-87)."""
    d = desc.contents if hasattr(desc, "contents") else desc
    pos = hp["position"].astype(np.float64)
    scale = float(np.abs(pos).max()) + 1e-30
    out = np.zeros(pos.shape[0], bool)
    for i in range(d.numLights):
        L = d.lights[i]
        if L.lightType != 0:
            continue
        n = np.array(list(L.normal), np.float64)
        n /= np.linalg.norm(n) + 1e-30
        out |= np.abs((pos - np.array(list(L.position), np.float64)) @ n) < 1e-4 * scale
    return out


def ppm_explained(a, b, radius):
    """Masks (per pixel) of the PPM rule after one iteration run on both renderers.  Returns dict(camera, indirect, direct)."""
    cam, hpa, hpb = camera_divergence(a, b)
    pa, pb = a.photons_as_emitted(), b.photons_as_emitted()
    va, vb = pa["power"].max(axis=1) > 0, pb["power"].max(axis=1) > 0
    scale = float(np.abs(pb["position"][vb]).max()) if vb.any() else 1.0
    # a photon "differs" when its position moved by more than 2e-4 r: the gather kernel's weight (IndirectRadianceEstimation.cu:59-67)
    # changes by 2.07 * delta(d^2 / r^2) of the photon's power, so a shift of 5e-4 r already moves a one-photon pixel by 2e-3
    close = np.isclose(pa["position"], pb["position"], rtol=0, atol=2e-4 * radius).all(axis=1) & \
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
    inputs = shading_inputs_differ(hpa, hpb) & diffuse
    thread = photon_thread_divergence(a, b, pa, pb) & diffuse
    return dict(hpa=hpa, hpb=hpb, camera=cam, indirect=cam | near | edge, direct=cam | (diffuse & mask_diff) | inputs | thread, inputs=inputs, thread=thread, photons_differ=int((~same).sum()),
                photons=int(vb.sum()), state=pixel_state_divergence(a, b), near=near, edge=edge, mask=mask_diff & diffuse)


# Gaussian filter of the gather (IndirectRadianceEstimation.cu:59-67) and the bound on its slope: |dw/dx| <= 2.07 for x = d^2 / r^2
_ALPHA, _BETA, _ENB = 1.818, 1.953, 0.141847


def _gather_weight(x):
    return _ALPHA * (1.0 - (1.0 - np.exp(-_BETA * x / 2.0)) / (1.0 - _ENB))


def _pair_contribution(hp, ph, pix, slot, valid, radius, emitted):
    """Contribution of photon `slot` to pixel `pix` (IndirectRadianceEstimation.cu:113-121, :213): the accept tests in float32 in
    the reference's operation order, the weight in float64."""
    f = np.float32
    diff = hp["position"][pix] - ph["position"][slot]
    d2 = (diff[:, 0] * diff[:, 0] + diff[:, 1] * diff[:, 1]) + diff[:, 2] * diff[:, 2]
    n, d = hp["normal"][pix], -ph["rayDirection"][slot]
    dot = (d[:, 0] * n[:, 0] + d[:, 1] * n[:, 1]) + d[:, 2] * n[:, 2]
    r2 = f(radius) * f(radius)
    acc = valid[slot] & (d2 <= r2) & (dot >= 0)
    w = _gather_weight(d2.astype(np.float64) / float(r2))
    c = hp["attenuation"][pix].astype(np.float64) / (np.pi * float(r2) * emitted)
    return np.where(acc[:, None], ph["power"][slot].astype(np.float64) * w[:, None] * c, 0.0)


def ppm_indirect_slack(a, b, radius, cam, hpa, hpb):
    """Quantitative form of the PPM indirect rule.  Photon slots are compared one by one: a slot DIFFERS when it is stored in only
    one map, or its position moved by more than 5e-5 r, or its power by more than 5e-5 (the slots that agree can together move a
    pixel by at most 2.07 * 2 * 5e-5 * sum(power) / sum(power * weight) ~ 4e-4 of its value, which the caller adds to the tolerance).
    For every (pixel, differing slot) pair in range the contribution is evaluated with each renderer's own photon and hit point,
    and |C_gpu - C_oracle| is summed per pixel: the SLACK by which the measured divergences of the photon maps can move that
    pixel.  Agreeing slots whose accept / reject is decided by ulps (within 2e-4 r of the radius sphere, or within 1e-4 of the
    tangent plane) add their full contribution.  Returns (slack (pixels, 3), stats dict)."""
    W, H = a.W, a.H
    emitted = float(a.pl[0] * a.pl[1])
    pa, pb = a.photons_as_emitted(), b.photons_as_emitted()
    va, vb = pa["power"].max(axis=1) > 0, pb["power"].max(axis=1) > 0
    dpos = np.abs(pa["position"] - pb["position"]).max(axis=1)
    dpow = np.abs(pa["power"] - pb["power"]).max(axis=1)
    differ = (va != vb) | (va & vb & ((dpos > 5e-5 * radius) | (dpow > 5e-5 * pb["power"].max(axis=1))))
    diffuse = ((hpb["flags"] & ol.PRD_HIT_NON_SPECULAR) != 0) & ~cam
    pix_ids = np.nonzero(diffuse)[0]
    slack = np.zeros((W * H, 3), np.float64)
    stats = dict(photons=int(vb.sum()), differ=int(differ.sum()), only_one=int((va != vb).sum()), pairs=0, edge_pairs=0)
    if pix_ids.size == 0:
        return slack, stats
    ptree = cKDTree(hpb["position"][pix_ids].astype(np.float64))

    def pairs_within(points, dist):
        if points.shape[0] == 0:
            return np.zeros(0, np.int64), np.zeros(0, np.int64)
        m = ptree.sparse_distance_matrix(cKDTree(points.astype(np.float64)), dist, output_type="ndarray")
        return m["i"].astype(np.int64), m["j"].astype(np.int64)

    # (1) differing slots: exact change of the contribution
    slots = np.nonzero(differ)[0]
    if slots.size:
        # every pixel within r of the slot's position in EITHER map (a diverged photon thread stores its photon somewhere else
        # entirely); the hit points of these pixels agree within `moved`, far below the 1e-3 r margin
        keys = []
        for ph, v in ((pa, va), (pb, vb)):
            sl = slots[v[slots]]
            i1, j1 = pairs_within(ph["position"][sl], radius * 1.001)
            keys.append(pix_ids[i1] * np.int64(pa.shape[0]) + sl[j1])
        key = np.unique(np.concatenate(keys))
        pix, slot = key // pa.shape[0], key % pa.shape[0]
        ca = _pair_contribution(hpa, pa, pix, slot, va, radius, emitted)
        cb = _pair_contribution(hpb, pb, pix, slot, vb, radius, emitted)
        np.add.at(slack, pix, np.abs(ca - cb))
        stats["pairs"] = int(pix.size)
    # (2) agreeing slots on the radius sphere or the tangent plane of a pixel: accept / reject decided by ulps
    same = np.nonzero(vb & ~differ)[0]
    if same.size:
        i2, j2 = pairs_within(pb["position"][same], radius * (1 + 2e-4))
        pix, slot = pix_ids[i2], same[j2]
        diff = (hpb["position"][pix] - pb["position"][slot]).astype(np.float64)
        d = np.sqrt((diff * diff).sum(axis=1))
        dot = -(pb["rayDirection"][slot].astype(np.float64) * hpb["normal"][pix]).sum(axis=1)
        dn = np.abs(hpa["normal"][pix].astype(np.float64) - hpb["normal"][pix]).max(axis=1)      # normal-mapped normals differ by the texture unit's 1/256 weights
        dh = np.abs(hpa["position"][pix].astype(np.float64) - hpb["position"][pix]).max(axis=1)
        sel = (np.abs(d - radius) <= 2e-4 * radius + 2.0 * dh) | (np.abs(dot) <= 1e-4 + 2.0 * dn)
        pix, slot = pix[sel], slot[sel]
        c = hpb["attenuation"][pix].astype(np.float64) / (np.pi * radius * radius * emitted)
        np.add.at(slack, pix, pb["power"][slot].astype(np.float64) * _ALPHA * c)
        stats["edge_pairs"] = int(pix.size)
    # (3) measured drift of the pixel's own inputs: a hit point that moved by dh shifts every distance by at most dh, which changes
    # sum(P w(d^2/r^2)) by at most 2.07 * 2 dh / r * sum(P) <= 8.33 dh / r of the sum (w >= 0.497 inside the radius); the attenuation
    # enters as a factor.  Added to the agreeing photons' 5e-5 r.  Returned as a RELATIVE allowance per pixel.
    dh = np.abs(hpa["position"].astype(np.float64) - hpb["position"]).max(axis=1)
    att = np.maximum(np.abs(hpb["attenuation"]).max(axis=1).astype(np.float64), 1e-30)
    datt = np.abs(hpa["attenuation"].astype(np.float64) - hpb["attenuation"]).max(axis=1) / att
    stats["drift"] = np.where(diffuse, 8.33 * (dh + 5e-5 * radius) / radius + datt, 0.0)
    return slack, stats


def within_tolerance_plus_slack(x, y, slack, drift, rtol=RTOL, atol_rel=ATOL_REL):
    """Per pixel: |x - y| <= atol + (rtol + drift) |y| + 1.01 slack on every channel; drift is the per-pixel relative allowance.
    Returns (ok mask, |x - y| / allowed per pixel)."""
    x, y = x.reshape(-1, 3).astype(np.float64), y.reshape(-1, 3).astype(np.float64)
    atol = atol_rel * float(np.abs(y).max() + 1e-30)
    allowed = atol + (rtol + np.asarray(drift, np.float64).reshape(-1, 1)) * np.abs(y) + 1.01 * slack
    ratio = (np.abs(x - y) / allowed).max(axis=1)
    return ratio <= 1.0, ratio


def gather_stage_exact(a, radius):
    """(A) of the PPM indirect rule: the product's gather against the ORACLE's gather run on the product's own inputs of the same
    iteration (its hit points, its sorted photon list, its offset table and grid): identical inputs, so EVERY pixel must agree
    within rel 1e-4 (ex2.approx and the summation order) — except, at most 4 per million, pixels that hold a photon whose accept /
    reject hangs on the last bit of distance2 or of the dot product, which must then differ by exactly that photon's contribution.
    Returns the worst relative error of the other pixels."""
    import ctypes as C
    r, W, H = a.r, a.W, a.H
    hp = a.hitpoints()
    g = r.debug_buffer("grid", np.uint32, 8)
    nv = int(g[7])
    gs = (C.c_uint32 * 3)(int(g[1]), int(g[2]), int(g[3]))
    gf = g.view(np.float32)
    org = (C.c_float * 3)(float(gf[4]), float(gf[5]), float(gf[6]))
    ncell = int(g[1]) * int(g[2]) * int(g[3])
    sorted_ph = r.debug_buffer("sorted_photons", ol.PHOTON_DT, max(nv, 1))
    offsets = r.debug_buffer("offsets", np.uint32, ncell + 1)
    out0 = np.zeros((W * H, 3), np.float32)
    rc = ol.load("oracle").ora_gather(W * H, ol.fptr(hp.view(np.float32)), nv, ol.fptr(sorted_ph.view(np.float32)), ol.uptr(offsets), float(gf[0]), gs, org,
                                      float(radius), float(a.pl[0] * a.pl[1]), ol.fptr(out0))
    assert rc == 0
    out = a.indirect()
    err = np.abs(out - out0) / (np.abs(out0) + 1e-6 * float(out0.max() + 1e-30))
    ok = np.isclose(out, out0, rtol=1e-4, atol=1e-6 * float(out0.max() + 1e-30)).all(axis=1)
    # The only admissible exception: a photon whose accept / reject hangs on the LAST BIT — the product evaluates the dot products as
    # one product + two FMAs (as nvcc compiles the reference's dot() for the GPU), the CPU oracle rounds every product.  Each such
    # pixel must hold a photon within 4 ulps of the radius sphere or of the tangent plane, and differ by that photon's contribution.
    bad = np.nonzero(~ok)[0]
    assert bad.size <= max(2, 4e-6 * W * H), "gather differs from the oracle's gather on identical inputs at %d pixels (worst relative error %.3e at pixel %d)" % (
        bad.size, err.max(), int(err.max(axis=1).argmax()))
    r2 = np.float32(radius) * np.float32(radius)
    for p in bad:
        pos, nrm = hp["position"][p], hp["normal"][p]
        diff = pos[None, :] - sorted_ph["position"]
        d2 = (diff[:, 0] * diff[:, 0] + diff[:, 1] * diff[:, 1]) + diff[:, 2] * diff[:, 2]
        nd = (sorted_ph["rayDirection"][:, 0] * nrm[0] + sorted_ph["rayDirection"][:, 1] * nrm[1]) + sorted_ph["rayDirection"][:, 2] * nrm[2]
        near = d2 <= r2 * np.float32(1.001)
        on_sphere = near & (np.abs(d2 / r2 - 1) <= 5e-7)
        on_plane = near & (np.abs(nd) <= 5e-7)
        cand = np.nonzero(on_sphere | on_plane)[0]
        assert cand.size, "gather differs from the oracle's gather on identical inputs at pixel %d (relative error %.3e) and no photon sits on its radius sphere or tangent plane" % (p, err[p].max())
        c = hp["attenuation"][p].astype(np.float64) / (np.pi * float(r2) * float(a.pl[0] * a.pl[1]))
        contrib = [sorted_ph["power"][j].astype(np.float64) * _gather_weight(float(d2[j]) / float(r2)) * c for j in cand]
        delta = np.abs(out[p].astype(np.float64) - out0[p])
        assert any(np.allclose(delta, x, rtol=2e-2, atol=1e-6 * float(out0.max())) for x in contrib), \
            "pixel %d: the difference %s is not the contribution of a photon on the radius sphere / tangent plane %s" % (p, delta, contrib)
    sizeable = ok & (np.abs(out0).max(axis=1) > 1e-3 * float(out0.max() + 1e-30))       # relative error is meaningless on pixels that are in by the absolute term
    return float(err[sizeable].max()) if sizeable.any() else 0.0


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
    (la, ca), (lb, cb) = a.path_signatures(), b.path_signatures()
    sig = (la != lb) | (ca != cb)              # the light or the camera subpath of this index met different materials
    state = state | sig
    count = a.connection_counts() != b.connection_counts()
    # same decisions, drifted geometry: a light subpath that stored the same number of vertices in both renderers but somewhere else (ulps
    # of sinf / cosf grow at every curved or normal-mapped bounce) changes every connection to it and moves its splats
    (patha, posa), (pathb, posb) = a.light_vertex_positions(), b.light_vertex_positions()
    oa, ob = np.argsort(patha, kind="stable"), np.argsort(pathb, kind="stable")     # path-major; a path's vertices keep their store order
    patha, posa, pathb, posb = patha[oa], posa[oa], pathb[ob], posb[ob]
    n = W * H
    ca, cb = np.bincount(patha, minlength=n), np.bincount(pathb, minlength=n)
    same_count = ca == cb
    ma, mb = same_count[patha], same_count[pathb]
    scale = float(np.abs(posb).max()) + 1e-30 if posb.size else 1.0
    moved = np.abs(posa[ma] - posb[mb]).max(axis=1) > 1e-5 * scale
    drift = np.bincount(patha[ma][moved], minlength=n) > 0
    state = state | drift | ~same_count
    splat = np.zeros(W * H, bool)
    for path, pos in ((patha, posa), (pathb, posb)):
        sel = state[path]
        _mark_projected(splat, W, H, *project_to_pixel(cam, W, H, pos[sel]))
    return dict(all=state | count | splat, state=state, count=count, splat=splat, sig=sig)


def path_tracing_explained(a, b):
    """Per-pixel mask of the PATH_TRACING rule: the payload RNG stream ended in a different state, or the path's signature — per
    segment: hit flags, light sample contributed or not, ended by Russian roulette (decisions RayGeneratorPT.cu:84-127 draws from
    the stream whose final state it discards) — differs."""
    div = pixel_state_divergence(a, b) | (a.path_signature() != b.path_signature())
    # same decisions, but the path itself drifted: ulps of sinf / cosf in a bounce direction grow at every curved or normal-mapped
    # surface, and after up to five bounces the last vertex sits elsewhere (another texel, another light distance)
    pa, pb = a.hitpoints()["position"], b.hitpoints()["position"]
    scale = float(np.abs(pb).max()) + 1e-30
    moved = np.abs(pa - pb).max(axis=1) > 1e-5 * scale
    return div | moved


def check_explained(name, bad, explained, extra="", allow=0):
    """Assert bad ⊂ explained (but for `allow` pixels: only the full-size cases pass a non-zero value, at most 4 per million, and print
    them); returns the one-line report the tests print."""
    n = bad.size
    unexplained = bad & ~explained
    line = "%s: %d px, out of tolerance %d (%.3f %%), of those explained %d, UNEXPLAINED %d; explained set covers %.2f %% of the image%s" % (
        name, n, int(bad.sum()), 100.0 * bad.mean(), int((bad & explained).sum()), int(unexplained.sum()), 100.0 * explained.mean(), extra)
    print(line)
    assert int(unexplained.sum()) <= allow, line + " first unexplained pixels: %s" % np.nonzero(unexplained)[0][:10].tolist()
    return line


def run_lockstep(name, r, o, det, odet, method, W, H, pl, radii, cam_pod, max_explained=0.5, per_iteration=None, allow=0):
    """Renders len(radii) iterations on the product `r` and the oracle `o` in lock-step and applies the rule to every iteration's
    buffers (PPM: indirect and direct; BDPT / PT: the increment of the sum buffer) and to the final sum buffer (explained = union
    over the iterations).  The explained set must stay below `max_explained` of the image, or the rule would explain anything.
    Returns dict(bad=fraction out of tolerance, explained=fraction explained, lines=[report lines])."""
    a, b = ProductProbe(r, W, H, pl), OracleProbe(o, W, H, pl)
    if method != 2:
        r.set_connection_counts(True)     # BDPT connection counts / PATH_TRACING path signatures
    union = np.zeros(W * H, bool)
    worst_explained = 0.0
    prev_a = prev_b = np.zeros((W * H, 3), np.float32)
    acc_a = None
    lines = []
    for i, radius in enumerate(radii):
        r.renderNextIteration(i, i, radius, True, det)
        o.render_next_iteration(i, i, radius, odet)
        if method == 2:
            e = ppm_explained(a, b, radius)
            worst = gather_stage_exact(a, radius)       # (A) same inputs -> every pixel, no exception
            # (B) end to end: in tolerance once the tolerance is widened by what the measured photon-map differences can change
            slack, st = ppm_indirect_slack(a, b, radius, e["camera"], e["hpa"], e["hpb"])
            ia, ib = a.indirect(), b.indirect()
            bad = out_of_tolerance(ia, ib)
            ok, ratio = within_tolerance_plus_slack(ia, ib, slack, st["drift"])
            worst_px = np.argsort(-np.where(e["camera"], 0.0, ratio))[:5]
            lines.append(check_explained("%s it%d indirect" % (name, i), ~ok, e["camera"],
                                         "; worst |diff| / allowed among pixels with identical camera paths: %s at pixels %s, median relative drift allowance %.1e"
                                         % (np.round(ratio[worst_px], 3).tolist(), worst_px.tolist(), float(np.median(st["drift"][st["drift"] > 0])) if (st["drift"] > 0).any() else 0.0) +
                                         "; out of the plain tolerance %.3f %%; gather vs oracle gather on the product's inputs: worst rel %.1e; photon slots differing %d of %d "
                                         "(%d in one map only), %d (pixel, differing photon) pairs, %d ulp-decided pairs; pixels whose slack exceeds their tolerance %.4f"
                                         % (100.0 * bad.mean(), worst, st["differ"], st["photons"], st["only_one"], st["pairs"], st["edge_pairs"],
                                            float((slack > ATOL_REL * float(np.abs(ib).max() + 1e-30) + RTOL * np.abs(ib)).any(axis=1).mean()))))
            e["indirect"] = e["camera"]       # the quantitative rule leaves only the camera-path divergences as (non-quantitative) explanations
            lines.append(check_explained("%s it%d direct" % (name, i), out_of_tolerance(a.direct(), b.direct()), e["direct"],
                                         "; occlusion mask differs %.4f, normal / attenuation of the hit differ %.4f, photon thread of the same RNG stream diverged %.4f"
                                         % (e["mask"].mean(), e["inputs"].mean(), e["thread"].mean())))
            union |= e["indirect"] | e["direct"]
            worst_explained = max(worst_explained, float((e["indirect"] | e["direct"]).mean()))
            # PPM_OUTPUT_PASS (ppm/Output.cu:32-37) on the product's own buffers: sum = direct + indirect, added to the running sum
            inc = a.direct() + ia
            acc_a = inc if acc_a is None else acc_a + inc
        else:
            if method == 1:
                e = bdpt_explained(a, b, cam_pod)
                ex, extra = e["all"], "; state %.4f, connection count %.4f, splat of a diverged path %.4f" % (e["state"].mean(), e["count"].mean(), e["splat"].mean())
            else:
                ex = path_tracing_explained(a, b)
                extra = "; state %.4f, path signature %.4f" % (pixel_state_divergence(a, b).mean(), (a.path_signature() != b.path_signature()).mean())
            oa, ob = a.output(), b.output()
            bad_inc = out_of_tolerance(oa - prev_a, ob - prev_b, rtol=2 * RTOL)
            for p in np.nonzero(bad_inc & ~ex)[0][:4]:    # what a person needs to name the difference, should the rule fail
                print("pixel %d increment %s / %s" % (p, (oa - prev_a)[p], (ob - prev_b)[p]))
            lines.append(check_explained("%s it%d increment" % (name, i), bad_inc, ex, extra, allow=allow))
            prev_a, prev_b = oa, ob
            union |= ex
            worst_explained = max(worst_explained, float(ex.mean()))
        if per_iteration is not None:
            per_iteration(i, a, b)
    oa, ob = a.output(), b.output()
    bad = out_of_tolerance(oa, ob, rtol=RTOL if acc_a is not None else 2 * RTOL)     # the same tolerance as the increments it is the sum of
    if acc_a is not None:
        # the running sum is exactly the sum of the per-iteration buffers checked above (float32, same order of additions)
        exact = float((acc_a.view(np.uint32) == oa.view(np.uint32)).all(axis=1).mean())
        lines.append("%s output accumulation over %d iterations: bit-identical to sum(direct + indirect) for %.4f of the pixels" % (name, len(radii), exact))
        print(lines[-1])
        assert np.allclose(acc_a, oa, rtol=2e-6, atol=0), "%s: output buffer is not the sum of the per-iteration direct + indirect buffers" % name
    if acc_a is None:     # BDPT / PATH_TRACING: the sum buffer under the union of the per-iteration explanations (PPM: proven above, term by term)
        lines.append(check_explained("%s sum buffer after %d iterations" % (name, len(radii)), bad, union, allow=allow * len(radii)))
    # a rule that explains most of the image explains nothing: the cap is on the largest per-iteration explained set (the union over
    # many iterations grows towards 1 by construction)
    assert worst_explained <= max_explained, "%s: the explained set of one iteration covers %.1f %% of the image" % (name, 100 * worst_explained)
    # whole-image relative L1 with both images clipped at the oracle's 99th percentile (a diverged path that reaches the light is a
    # firefly of a two-iteration render and would otherwise carry the sum)
    clip = float(np.percentile(ob, 99.0)) + 1e-30
    rel = float(np.abs(np.minimum(oa, clip) - np.minimum(ob, clip)).sum() / (np.abs(np.minimum(ob, clip)).sum() + 1e-30))
    return dict(bad=float(bad.mean()), explained=float(union.mean()), rel_l1=rel, lines=lines)
