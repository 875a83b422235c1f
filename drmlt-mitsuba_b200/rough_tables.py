"""rough_tables.py -- the rough Fresnel transmittance tables of `roughplastic` for hosts that are not Mitsuba.

The reference precomputes the transmittance through a rough dielectric boundary over (eta, alpha, cos theta) and ships it as
data/microfacet/{beckmann,ggx}.dat; RoughPlastic::configure reduces it to the material's (eta, alpha) with tricubic
interpolation (src/bsdfs/rtrans.h:205-347, src/libcore/spline.cpp:236-304 and its 3-D sibling).  `reduce` restates that
reduction on the caller's copy of the .dat file and returns the DR_ROUGH_TABLE_DOUBLES doubles that
dr_scene_desc.rough_tables carries per roughplastic material (include/drmlt_b200.h).  A Mitsuba host does not need this
module: the plugin shim asks the reference's own RoughTransmittance (shim/mts_plugin.cpp).

Pinned against the reference's class by tests/test_ref_pins.py (tests/golden/ref_rough_tables.npz)."""
import struct

import numpy as np

from . import abi

HEADER = b"MTS_TRANSMITTANCE"


def load(path):
    """-> dict(trans [2*nEta, nAlpha, nTheta], diff [2*nEta, nAlpha], eta_min, eta_max, alpha_min, alpha_max) (rtrans.h:36-105)."""
    with open(path, "rb") as f:
        raw = f.read()
    if raw[:len(HEADER)] != HEADER:
        raise ValueError("Encountered an invalid transmittance data file!")
    o = len(HEADER)
    n_eta, n_alpha, n_theta = struct.unpack_from("<QQQ", raw, o)
    o += 24
    eta_min, eta_max, alpha_min, alpha_max = [float(np.float32(v)) for v in struct.unpack_from("<4f", raw, o)]
    o += 16
    data = np.frombuffer(raw, dtype="<f4", offset=o).astype(np.float64)
    if data.size != 2 * n_eta * n_alpha * (n_theta + 1):
        raise ValueError("transmittance data file has the wrong size")
    data = data.reshape(2 * n_eta, n_alpha, n_theta + 1)
    return dict(trans=np.ascontiguousarray(data[..., :n_theta]), diff=np.ascontiguousarray(data[..., n_theta]),
                eta_min=eta_min, eta_max=eta_max, alpha_min=alpha_min, alpha_max=alpha_max)


def _weights(x, size):
    """knot index and the four node weights of evalCubicInterp*D along one dimension (spline.cpp:242-287), x in [0, 1]."""
    if not (0.0 <= x <= 1.0):
        return None
    t = x * (size - 1) / 1.0
    k = min(int(t), size - 2)
    t -= k
    t2, t3 = t * t, t * t * t
    w = [0.0, 2 * t3 - 3 * t2 + 1, -2 * t3 + 3 * t2, 0.0]
    d0, d1 = t3 - 2 * t2 + t, t3 - t2
    if k > 0:
        w[2] += 0.5 * d0; w[0] -= 0.5 * d0
    else:
        w[2] += d0; w[1] -= d0
    if k + 2 < size:
        w[3] += 0.5 * d1; w[1] -= 0.5 * d1
    else:
        w[2] += d1; w[1] -= d1
    return k, w


def interp(values, point):
    """evalCubicInterp{1,2,3}D(point, values, values.shape reversed, 0, 1): `point` is ordered fastest-varying axis first,
    as the reference's Point2 / Point3 arguments are; `values` is the C-ordered array."""
    dims = len(point)
    kw = [_weights(point[d], values.shape[dims - 1 - d]) for d in range(dims)]
    if any(k is None for k in kw):
        return 0.0
    result = 0.0
    for offs in np.ndindex(*([4] * dims)):            # offs[0] = slowest axis, as the reference's outer loop
        w = 1.0
        for a in range(dims):                          # a-th array axis <-> point component dims - 1 - a
            w *= kw[dims - 1 - a][1][offs[a]]
        if w == 0.0:
            continue
        idx = tuple(kw[dims - 1 - a][0] + offs[a] - 1 for a in range(dims))
        result += values[idx] * w
    return result


def _warp(v, lo, hi):
    return ((v - lo) / (hi - lo)) ** 0.25


def _slice_eta(tab, eta):
    """RoughTransmittance::setEta (rtrans.h:205-262): -> (trans [nAlpha, nTheta], diff [nAlpha])."""
    n_eta = tab["trans"].shape[0] // 2
    trans, diff = tab["trans"][:n_eta], tab["diff"][:n_eta]
    if eta < 1:
        trans, diff = tab["trans"][n_eta:], tab["diff"][n_eta:]
        eta = 1.0 / eta
    eta = max(eta, tab["eta_min"])
    we = _warp(eta, tab["eta_min"], tab["eta_max"])
    n_alpha, n_theta = trans.shape[1], trans.shape[2]
    # `Float dAlpha = 1.0f / (m_alphaSamples - 1)`: a FLOAT division (rtrans.h:232-233), so the knots sit at i * float(1 / 49)
    d_alpha, d_theta = float(np.float32(1.0) / np.float32(n_alpha - 1)), float(np.float32(1.0) / np.float32(n_theta - 1))
    new_trans = np.array([[interp(trans, (j * d_theta, i * d_alpha, we)) for j in range(n_theta)] for i in range(n_alpha)])
    new_diff = np.array([interp(diff, (i * d_alpha, we)) for i in range(n_alpha)])
    return new_trans, new_diff


def reduce(path, eta, alpha):
    """The table of one roughplastic material: distribution file `path`, eta = intIOR / extIOR, alpha (the value the material
    carries; the reference averages its constant texture with a float third first, spectrum.h:481-486 -- done here too)."""
    tab = load(path) if not isinstance(path, dict) else path
    eta = float(np.float32(eta))                       # dr_material.eta[0] / .alpha are floats
    a = float(np.float32(alpha))
    a = (a + a + a) * float(np.float32(1.0) / np.float32(3.0))
    e = eta if eta >= 1 else 1.0 / eta
    if not (tab["eta_min"] <= e <= tab["eta_max"]):
        raise ValueError("the requested relative index of refraction eta=%f is outside of the supported range [%f, %f]" % (eta, tab["eta_min"], tab["eta_max"]))
    if not (tab["alpha_min"] <= a <= tab["alpha_max"]):
        raise ValueError("the requested roughness value alpha=%f is outside of the supported range [%f, %f]" % (a, tab["alpha_min"], tab["alpha_max"]))
    wa = _warp(a, tab["alpha_min"], tab["alpha_max"])
    ext_trans, ext_diff = _slice_eta(tab, eta)
    _, int_diff = _slice_eta(tab, 1.0 / eta)
    n_theta = ext_trans.shape[1]
    if n_theta != abi.DR_ROUGH_TABLE_THETA:
        raise ValueError("unexpected number of theta samples")
    out = np.zeros(abi.DR_ROUGH_TABLE_DOUBLES)
    d_theta = float(np.float32(1.0) / np.float32(n_theta - 1))     # (float division again, rtrans.h:329)
    for i in range(n_theta):                           # setAlpha (rtrans.h:331-340)
        out[i] = interp(ext_trans, (i * d_theta, wa))
    clamp = lambda v: min(1.0, max(0.0, v))
    out[100] = clamp(interp(int_diff, (wa,)))          # m_internalRoughTransmittance->evalDiffuse(alpha) (rtrans.h:177-183)
    out[101] = clamp(interp(ext_diff, (wa,)))
    return out
