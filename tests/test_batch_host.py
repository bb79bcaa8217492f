"""Host logic of the batched-systems upload (olb_prep.h::prepare_batch, the body of
olb_table_upload_batch) on the CPU: the prepared bytes of system b must equal the prepared bytes of the
single-system table that ``batch.system_table`` describes (only the feature word of the header may differ: it
carries the union over the batch), and structure-changing parameters must be rejected."""
import ctypes as C

import numpy as np
import pytest

from optiland_b200 import _lib
from optiland_b200 import table as T
from optiland_b200.batch import system_table, template_params
from tests._util import Case
from tests.test_hostcheck import hc  # noqa: F401  (fixture)

FEATURES_OFFSET = 12   # PrepHeader {n_surf, n_wl, pool_len, features, ...}


def _batch_blob(hc, table, P, b, which):
    ht = _lib.HostTable(table)
    buf = (C.c_ubyte * (1 << 20))()
    feat = C.c_uint(0)
    err = C.create_string_buffer(256)
    Pc = np.ascontiguousarray(P, dtype=np.float64)
    n = hc.olbhc_batch_blob(C.byref(ht.c), C.c_void_p(Pc.ctypes.data), P.shape[0], b, which, buf, len(buf), C.byref(feat), err, 256)
    assert n > 0, err.value
    return bytes(buf[:n]), feat.value


def _single_blob(hc, table, which):
    ht = _lib.HostTable(table)
    buf = (C.c_ubyte * (1 << 20))()
    feat = C.c_uint(0)
    err = C.create_string_buffer(256)
    n = hc.olbhc_single_blob(C.byref(ht.c), which, buf, len(buf), C.byref(feat), err, 256)
    assert n > 0, err.value
    return bytes(buf[:n]), feat.value


@pytest.mark.parametrize("name", ["cooke_c1", "aspheric_singlet", "tilted_fold", "hubble_c4"])
def test_batch_blobs_equal_single_system_blobs(hc, name):
    c = Case(name)
    table = c.table
    rng = np.random.default_rng(7)
    B = 4
    p0 = template_params(table)
    P = np.repeat(p0[None], B, axis=0)
    for b in range(1, B):
        for s, spec in enumerate(table.surfaces):
            if spec.kind == T.GEOM_NOOP:
                continue
            P[b, s, _lib.BP_TX:_lib.BP_TX + 3] += rng.normal(0, 0.02, 3)
            R = T.rotation_matrix(*rng.normal(0, 2e-3, 3))
            P[b, s, _lib.BP_R:_lib.BP_R + 9] = (R @ p0[s, _lib.BP_R:_lib.BP_R + 9].reshape(3, 3)).reshape(9)
            if spec.kind != T.GEOM_PLANE:
                P[b, s, _lib.BP_CURV] *= 1 + rng.normal(0, 1e-3)
                P[b, s, _lib.BP_CONIC] += rng.normal(0, 1e-3)
            if spec.kind == T.GEOM_EVEN_ASPHERE:
                k = len(spec.coefficients)
                P[b, s, _lib.BP_COEF:_lib.BP_COEF + k] *= 1 + rng.normal(0, 1e-2, k)
            P[b, s, _lib.BP_N1] += 1e-4 * (P[b, s, _lib.BP_N1] != 1.0)
    for which in (0, 1):
        union = 0
        singles = []
        for b in range(B):
            blob, feat = _single_blob(hc, system_table(table, P[b]), which)
            singles.append(blob)
            union |= feat
        for b in range(B):
            got, feat = _batch_blob(hc, table, P, b, which)
            assert feat == union
            want = bytearray(singles[b])
            want[FEATURES_OFFSET:FEATURES_OFFSET + 4] = int(union).to_bytes(4, "little")
            assert got == bytes(want), (name, which, b)


def test_batch_rejects_bad_parameters(hc):
    c = Case("cooke_c1")
    P = np.repeat(template_params(c.table)[None], 2, axis=0)
    ht = _lib.HostTable(c.table)
    buf = (C.c_ubyte * (1 << 20))()
    feat = C.c_uint(0)
    err = C.create_string_buffer(256)
    bad = P.copy()
    bad[1, 2, _lib.BP_R] = np.nan
    assert hc.olbhc_batch_blob(C.byref(ht.c), C.c_void_p(bad.ctypes.data), 2, 0, 0, buf, len(buf), C.byref(feat), err, 256) == -1
    assert b"system 1" in err.value and b"pose" in err.value
    m = Case("dgauss_multiwl")
    htm = _lib.HostTable(m.table)
    Pm = np.zeros((1, m.table.num_surfaces, _lib.BP_COUNT))
    assert hc.olbhc_batch_blob(C.byref(htm.c), C.c_void_p(Pm.ctypes.data), 1, 0, 0, buf, len(buf), C.byref(feat), err, 256) == -1
    assert b"one wavelength" in err.value
