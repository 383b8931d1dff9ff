"""Self-test of the parity rule (tests/parity_rules.py) on the CPU: two ORACLE renders whose cameras differ by a few ulps stand in
for the GPU-vs-CPU pair — float results differ slightly everywhere, a few paths flip a discrete decision.  Every pixel that leaves
the tolerance must be explained by a measured divergence, and a deliberately corrupted image must NOT be explained."""
import numpy as np
import pytest

import oracle_lib as ol
import parity_rules as pr
from oppositerenderer_b200 import capi


def _pair(oracle, method, W, H, pl, iters, r0_scale=1.0):
    sc = oracle.BuiltinScene("Cornell", 0)
    cam = sc.camera(W / H)
    cam2 = capi.opr_camera.from_buffer_copy(bytes(memoryview(cam).cast("B")))
    cam2.eye[0] = float(np.float32(cam2.eye[0]) * np.float32(1 + 3e-7)) + 1e-6
    A = oracle.OracleRenderer("oracle", pl, 8, 1645301512)
    B = oracle.OracleRenderer("oracle", pl, 8, 1645301512)
    A.init_scene(sc.desc)
    B.init_scene(sc.desc)
    radii = oracle.ppm_radius_schedule(sc.initial_radius() * r0_scale, iters)
    return sc, cam, cam2, A, B, radii


@pytest.mark.parametrize("method", [2, 1])
def test_rule_explains_ulp_level_divergence(oracle, method):
    W, H, pl, iters = 48, 48, (48, 48), 3
    sc, cam, cam2, A, B, radii = _pair(oracle, method, W, H, pl, iters)
    a, b = pr.OracleProbe(A, W, H, pl), pr.OracleProbe(B, W, H, pl)
    union = np.zeros(W * H, bool)
    for i in range(iters):
        A.render_next_iteration(i, i, radii[i], oracle.make_details(cam2, method, W, H))
        B.render_next_iteration(i, i, radii[i], oracle.make_details(cam, method, W, H))
        if method == 2:
            e = pr.ppm_explained(a, b, radii[i])
            pr.check_explained("indirect", pr.out_of_tolerance(a.indirect(), b.indirect()), e["indirect"])
            pr.check_explained("direct", pr.out_of_tolerance(a.direct(), b.direct()), e["direct"])
            union |= e["indirect"] | e["direct"]
        else:
            union |= pr.bdpt_explained(a, b, cam)["all"]
    pr.check_explained("sum", pr.out_of_tolerance(a.output(), b.output()), union)
    assert union.mean() < 0.5


def test_rule_rejects_an_unexplained_difference(oracle):
    W, H, pl = 48, 48, (48, 48)
    sc, cam, cam2, A, B, radii = _pair(oracle, 2, W, H, pl, 1)
    a, b = pr.OracleProbe(A, W, H, pl), pr.OracleProbe(B, W, H, pl)
    A.render_next_iteration(0, 0, radii[0], oracle.make_details(cam2, 2, W, H))
    B.render_next_iteration(0, 0, radii[0], oracle.make_details(cam, 2, W, H))
    e = pr.ppm_explained(a, b, radii[0])
    wrong = a.indirect().copy()
    victim = int(np.nonzero(~e["indirect"] & (b.indirect().sum(axis=1) > 0))[0][0])
    wrong[victim] *= 1.01            # a 1 % error on a pixel nothing explains
    with pytest.raises(AssertionError, match="UNEXPLAINED 1"):
        pr.check_explained("corrupted", pr.out_of_tolerance(wrong, b.indirect()), e["indirect"])
