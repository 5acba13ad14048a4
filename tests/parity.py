"""Shared parity metrics: how a traced image is compared with the oracle's (BASELINE.json north_star)."""
import numpy as np

ID_AGREEMENT_MIN = 0.999   # terminating object id must match on >= 99.9 % of pixels
STATE_RTOL = 1e-8          # final position and momentum, relative to the ray's own inf-norm
RGB_ATOL = 1.0 / 255.0     # colour tolerance


def state_rel_err(fs_ref, fs_out):
    """Per-ray relative error of position (x^0..x^3) and momentum (u^0..u^3), each relative to the
    inf-norm of that 4-vector in the reference."""
    ex = np.abs(fs_ref[:, :4] - fs_out[:, :4]).max(axis=1) / np.abs(fs_ref[:, :4]).max(axis=1)
    eu = np.abs(fs_ref[:, 4:] - fs_out[:, 4:]).max(axis=1) / np.abs(fs_ref[:, 4:]).max(axis=1)
    return ex, eu


def rgb_err(rgb_ref, rgb_out):
    """Channel-wise error; the first two sphere channels are mod(.,1) patterns, so a value just
    below 1 and one just above 0 are neighbours: compare on the circle."""
    d = np.abs(rgb_ref - rgb_out)
    return d


def compare(ref, out, rgb_ref, rgb_out, label=""):
    """Returns a dict of parity numbers and asserts the north-star bars."""
    same = ref["obj_id"] == out["obj_id"]
    ex, eu = state_rel_err(ref["final_state"], out["final_state"])
    ok_state = (ex <= STATE_RTOL) & (eu <= STATE_RTOL)
    d = rgb_err(rgb_ref, rgb_out).max(axis=1)
    res = dict(n=int(same.size), id_agree=float(same.mean()), n_id_mismatch=int((~same).sum()),
               state_ok=float(ok_state[same].mean()) if same.any() else 1.0,
               n_state_bad=int((~ok_state[same]).sum()),
               max_ex=float(ex[same].max()), max_eu=float(eu[same].max()),
               rgb_max=float(d[same].max()), n_rgb_bad=int((d[same] > RGB_ATOL).sum()))
    return res


def ks_metric_numpy(x, M, a):
    """g_ab = eta_ab + f k_a k_b of src:274-294 with the radius line src:284 AS WRITTEN, for an (n, >=4) array of
    positions -- an independent numpy evaluation (no derivative, no Christoffel symbol) for the conservation laws."""
    X, Y, Z = x[:, 1], x[:, 2], x[:, 3]
    rho2 = X * X + Y * Y + Z * Z
    r = np.sqrt(rho2 - a * a) / 2 + np.sqrt(a * a * Z * Z + ((rho2 - a * a) / 2) ** 2)
    f = 2 * M * r ** 3 / (r ** 4 + a * a * Z * Z)
    k = np.stack([np.ones_like(r), (r * X + a * Y) / (r * r + a * a), (r * Y - a * X) / (r * r + a * a), Z / r], axis=1)
    g = np.zeros((len(r), 4, 4))
    g[:, 0, 0] = -1
    g[:, 1, 1] = g[:, 2, 2] = g[:, 3, 3] = 1
    return g + f[:, None, None] * k[:, :, None] * k[:, None, :]


def conservation_errors(s0, s1, M, a, chunk=200000):
    """Per ray: |g(u,u)| at the final state s1 relative to sum |g_ab u^a u^b|, and the relative drift of the
    Killing energy g_{0b} u^b between the initial state s0 and s1 (the metric is stationary)."""
    null_rel = np.empty(len(s1))
    e_rel = np.empty(len(s1))
    for lo in range(0, len(s1), chunk):
        a0, a1 = s0[lo:lo + chunk], s1[lo:lo + chunk]
        g0, g1 = ks_metric_numpy(a0, M, a), ks_metric_numpy(a1, M, a)
        u0, u1 = a0[:, 4:8], a1[:, 4:8]
        null_rel[lo:lo + chunk] = np.abs(np.einsum("nab,na,nb->n", g1, u1, u1)) / \
            np.einsum("nab,na,nb->n", np.abs(g1), np.abs(u1), np.abs(u1))
        e0 = np.einsum("nb,nb->n", g0[:, 0, :], u0)
        e1 = np.einsum("nb,nb->n", g1[:, 0, :], u1)
        e_rel[lo:lo + chunk] = np.abs(e1 - e0) / np.abs(e0)
    return null_rel, e_rel
