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
