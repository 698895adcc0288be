"""CPU-side checks of the product: the C-ABI library loads and exports every symbol the header declares,
fails loudly without a GPU, and the host-side mirror logic (term packing, Hilbert encodings, conversions)
agrees with the oracle.  No compute calls."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch
from conftest import ROOT, load_table, load_terms_json

import naqs_b200
from naqs_b200 import _lib
from oracle import eloc_oracle as eo

HAS_GPU = torch.cuda.is_available()


def header_symbols():
    text = open(os.path.join(ROOT, "include", "naqs_eloc.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(naqs_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(built_library):
    lib = ctypes.CDLL(built_library)
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/naqs_eloc.h but not exported"
    assert set(syms) == set(_lib.SIGNATURES), "ctypes signature table and header disagree"
    assert _lib.load().naqs_abi_version() == 1


def test_header_cites_reference_lines():
    text = open(os.path.join(ROOT, "include", "naqs_eloc.h")).read()
    for cite in ("hamiltonian_math.pyx:200-288", "hamiltonian_math.pyx:455-484", "sparse_math.pyx:49-243", "sparse_math.pyx:251-402",
                 "hilbert_math.pyx:12-44", "hamiltonian.py:272-370", "energy.py:219-263", "hilbert.py:607-640"):
        assert cite in text


def test_argument_errors_need_no_gpu():
    lib = _lib.load()
    # dtype errors are raised before any device work (hamiltonian_math.pyx:484 -> TypeError)
    with pytest.raises(TypeError):
        _lib.check(lib.naqs_popcount_parity(None, 3, 10, None, None))
    with pytest.raises(TypeError):
        _lib.check(lib.naqs_sparse_dense_mv(None, 2, None, None, 4, 1, None, None, None))
    with pytest.raises(ValueError):
        out = ctypes.c_void_p()
        _lib.check(lib.naqs_table_create(ctypes.byref(out), None, None, None, 0, 3, 10, -1, -1, 0))
    with pytest.raises(ValueError):
        out = ctypes.c_void_p()
        _lib.check(lib.naqs_table_create(ctypes.byref(out), None, None, None, 0, 1, 64, -1, -1, 0))


@pytest.mark.skipif(HAS_GPU, reason="checks the no-GPU failure mode")
def test_fails_loudly_without_gpu():
    xy, yz, c, N, na, nb = load_table("LiH")
    with pytest.raises(naqs_b200.NaqsError):
        naqs_b200.DeviceTermTable(xy, yz, c, N, na, nb)
    with pytest.raises(naqs_b200.NaqsError):
        naqs_b200.hamiltonian_math.popcount_parity(np.arange(5, dtype=np.int32))
    with pytest.raises(naqs_b200.NaqsError):
        naqs_b200.hilbert_math.make_basis_idxs_cy(3)
    out = ctypes.c_void_p()
    rc = _lib.load().naqs_table_create(ctypes.byref(out), None, None, None, 0, 1, 10, -1, -1, 0)
    assert rc == 3 and b"no CPU fallback" in _lib.load().naqs_last_error()  # NAQS_ERR_CUDA


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "naqs-for-quantum-chemistry_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f"{f} imports the oracle"
                assert "liboracle" not in text and "eloc_oracle" not in text


@pytest.mark.parametrize("mol", ["H2", "LiH"])
def test_pack_terms_product_matches_reference_table(mol):
    xy, yz, c, N, na, nb = load_table(mol)
    pxy, pyz, pc = naqs_b200.pack_terms(load_terms_json(mol), N)
    assert np.array_equal(pxy[:, 0], xy) and np.array_equal(pyz[:, 0], yz) and np.array_equal(pc, c)


@pytest.mark.skipif(not os.path.isdir("/root/reference/molecules"), reason="needs the reference's pickles (build container only)")
@pytest.mark.parametrize("mol", ["H2O", "N2", "N2_1.5", "H2S", "NH3", "C2", "Li2O"])
def test_pkl_loader_and_pack_terms_reproduce_reference_tables_bitwise(mol):
    """load_qubit_hamiltonian (openfermion-free unpickler) + pack_terms against the tables the reference's own
    __calc_coupling_info produced (tests/golden/make_golden.py), bit patterns included: H2O / N2_1.5 / H2S carry odd-nY
    terms whose coefficients are signed zeros (hamiltonian.py:416,424)."""
    xy, yz, c, N, na, nb = load_table(mol)
    op = naqs_b200.load_qubit_hamiltonian(f"/root/reference/molecules/{mol}/{mol}_qubit_hamiltonian.pkl")
    assert op.many_body_order() <= N
    pxy, pyz, pc = naqs_b200.pack_terms(op.terms, N)
    assert pxy.shape == (len(c), 1)
    assert np.array_equal(pxy[:, 0], xy) and np.array_equal(pyz[:, 0], yz)
    assert np.array_equal(pc.view(np.uint64), c.view(np.uint64))  # bitwise: -0.0 != +0.0 here
    if mol in ("H2O", "N2_1.5", "H2S"):
        n_y = np.array([bin(int(a) & int(b)).count("1") for a, b in zip(pxy[:, 0], pyz[:, 0])])
        assert (n_y % 2 == 1).any() and np.all(pc[n_y % 2 == 1] == 0.0)


def test_pack_terms_filters_and_odd_y():
    terms = {(): 1.5 + 0j, ((0, "X"), (1, "Y")): 2.0 + 0j, ((0, "Y"), (3, "Y")): 0.25 + 0j, ((2, "Z"),): -1.0 + 0j,
             ((1, "X"), (2, "X"), (3, "X"), (4, "X")): 0.5 + 1e-9j}
    xy, yz, c = naqs_b200.pack_terms(terms, 6)
    oxy, oyz, oc = eo.pack_terms(terms, 6)
    assert np.array_equal(xy, oxy) and np.array_equal(yz, oyz) and np.array_equal(c, oc)
    assert c[1] == 0.0 and c[2] == -0.25            # odd-nY zeroed, i^2 = -1 (hamiltonian.py:416)
    xy2, _, c2 = naqs_b200.pack_terms(terms, 6, n_occ=1)          # terms flipping qubit 0 dropped (:394-396)
    assert len(c2) == 3
    xy3, _, c3 = naqs_b200.pack_terms(terms, 6, n_excitations_max=2)  # 4-flip term dropped (:397-401)
    assert len(c3) == 4
    for kw in ({"n_occ": 1}, {"n_excitations_max": 2}):
        a, b = naqs_b200.pack_terms(terms, 6, **kw), eo.pack_terms(terms, 6, **kw)
        assert all(np.array_equal(x, y) for x, y in zip(a, b))


def test_pack_terms_wide():
    terms = {((70, "X"), (3, "Y"), (100, "Z")): 1.0 + 0j, ((126, "Z"),): 2.0 + 0j}
    xy, yz, c = naqs_b200.pack_terms(terms, 127)
    oxy, oyz, oc = eo.pack_terms(terms, 127)
    assert xy.shape == (2, 2) and np.array_equal(xy, oxy) and np.array_equal(yz, oyz) and np.array_equal(c, oc)


@pytest.mark.parametrize("shape", [(4, 1, 1), (12, 2, 2), (14, 5, 5), (9, 2, 3), (16, 5, 5)])
def test_hilbert_mirror_matches_oracle(shape):
    N, na, nb = shape
    h = naqs_b200.Hilbert.get(N, na, nb, encoding=naqs_b200.Encoding.SIGNED)
    sec = eo.sector_keys(N, na, nb)[:, 0].astype(np.int64)
    assert h.size == len(sec)
    assert np.array_equal(h.sector_keys(), sec)
    idx = h.get_subspace(ret_states=False, ret_idxs=True)
    assert idx.dtype == h._idx_torch_dtype and np.array_equal(idx.numpy().astype(np.int64), sec)
    assert np.array_equal(np.asarray(h.full2restricted_idx(sec.astype(h._idx_np_dtype))).astype(np.int64), np.arange(len(sec)))
    allk = np.arange(2 ** N, dtype=np.int64)
    r = np.asarray(h.full2restricted_idx(allk.astype(h._idx_np_dtype))).astype(np.int64)
    assert np.array_equal(r, eo.restricted_index(eo.as_keys(allk), N, na, nb))
    states = h.idx2state(sec[:20].astype(h._idx_np_dtype))
    assert states.dtype == torch.int8 and set(np.unique(states.numpy())) <= {-1, 1}
    assert np.array_equal(h.state2idx(states).numpy().reshape(-1).astype(np.int64), sec[:20])
    n5 = min(5, len(sec))
    assert np.array_equal(np.asarray(h.restricted2full_idx(np.arange(n5))).astype(np.int64), sec[:n5])


def test_idx_dtype_rule():
    for N, dt in [(12, np.int16), (15, np.int16), (16, np.int32), (29, np.int32), (30, np.int64)]:
        assert naqs_b200.Hilbert.get(N, encoding=naqs_b200.Encoding.SIGNED)._idx_np_dtype is dt is eo.idx_dtype(N)


def test_key_and_psi_conversions():
    k = _lib.keys_to_numpy(np.array([1, 5, -1], np.int16), 1)
    assert k.dtype == np.uint64 and k.shape == (3, 1) and int(k[2, 0]) == 2 ** 64 - 1
    k2 = _lib.keys_to_numpy(np.array([1 << 70, 3], dtype=object), 2)
    assert k2.shape == (2, 2) and int(k2[0, 1]) == 1 << 6 and int(k2[1, 0]) == 3
    with pytest.raises(TypeError):
        _lib.keys_to_numpy(np.array([1.5]), 1)
    s = naqs_b200.stats_from_sums(np.array([2.0, -3.0, 1.0, 5.0, 7.0]))
    assert s["mean"] == complex(-1.5, 0.5) and abs(s["variance"] - 0.25) < 1e-15 and s["n"] == 7


def test_type_mv_promotion_rules():
    from scipy.sparse import identity
    from naqs_b200 import sparse_math
    m64, m32 = identity(3, dtype=np.float64, format="csr"), identity(3, dtype=np.float32, format="csr")
    cases = [(m32, np.complex64, 32, np.complex64), (m32, np.complex128, 64, np.complex128), (m64, np.complex64, 64, np.complex128),
             (m64, np.float64, 64, np.complex128), (m32, np.float32, 32, np.complex64)]
    for m, vdt, bits, out_dt in cases:
        mm, vv, nb, _ = sparse_math._type_mv(m, np.ones(3, vdt))
        assert nb == bits and vv.dtype == out_dt and mm.dtype == (np.float64 if bits == 64 else np.float32)
    with pytest.raises(Exception):
        sparse_math._type_mv(identity(3, dtype=np.int32, format="csr"), np.ones(3))


@pytest.mark.skipif(not os.path.isdir("/root/reference/src/optimizer"), reason="reference tree not present (GPU box)")
def test_install_patches_the_unmodified_reference():
    """install() makes the reference's own import statements resolve to the device-backed twins and rebinds the two
    Level-1 seams; run in a subprocess so the patched modules do not leak into the other tests."""
    import subprocess
    import sys
    code = r'''
import sys
sys.path.insert(0, "%s")
from oracle import ref_harness
ref_harness.install_stubs()            # openfermion / torch._six stand-ins the reference needs to import at all
import naqs_b200
assert naqs_b200.install("/root/reference")
import src.utils.hamiltonian_math as hm, src.utils.sparse_math as sm, src.utils.hilbert_math as him
assert hm is naqs_b200.hamiltonian_math and sm is naqs_b200.sparse_math and him is naqs_b200.hilbert_math
import src.optimizer.hamiltonian as H, src.optimizer.energy as E
assert H.get_Hij_cy is naqs_b200.hamiltonian_math.get_Hij_cy          # hamiltonian.py:13 picked up the twin
assert E.sparse_dense_mv is naqs_b200.sparse_math.sparse_dense_mv    # energy.py:27
assert E.OptimizerBase.calculate_local_energy is naqs_b200.calculate_local_energy
assert H.PauliHamiltonian.get is naqs_b200.PauliHamiltonian.get
import inspect
ref_sig = ["self", "states_idx", "psi", "set_unsampled_states_to_zero", "ret_complex"]
params = inspect.signature(naqs_b200.calculate_local_energy).parameters
assert list(params)[:len(ref_sig)] == ref_sig                          # the reference's arguments, same order (energy.py:220)
assert all(params[k].default is not inspect.Parameter.empty for k in list(params)[len(ref_sig):])  # extras are optional
print("ok")
''' % ROOT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr


@pytest.mark.skipif(not os.path.isdir("/root/reference/src/optimizer"), reason="reference tree not present (GPU box)")
def test_install_device_resident_patches_every_hilbert_flavour():
    """install(device_resident=True): state2idx is defined per Hilbert flavour (hilbert.py:344, 573, 833) and to_idx_array once in
    the base class (:85) — every concrete state2idx and the base to_idx_array must be wrapped (a CUDA tensor reaching an unwrapped
    one fails inside the reference with a device mismatch), host inputs must still take the reference's own code, and
    _SGD_step must be the fused step."""
    import subprocess
    import sys
    code = r'''
import sys, types
sys.path.insert(0, "%s")
from oracle import ref_harness
ref_harness.install_stubs()
import numpy as np, torch
import naqs_b200
assert naqs_b200.install("/root/reference")                                  # the twins make `import src.utils.hilbert` possible
import src.utils.hilbert as hil
orig = {n: getattr(hil, n).__dict__["state2idx"] for n in ("_HilbertFull", "_HilbertRestricted", "_HilbertPartiallyRestricted")}
orig_arr = hil._HilbertBase.__dict__["to_idx_array"]
assert naqs_b200.install("/root/reference", device_resident=True)
for n, f in orig.items():
    g = getattr(hil, n).__dict__["state2idx"]
    assert g is not f and getattr(g, "__wrapped__", None) is f, n          # wrapped, and the host path is the reference's own function
assert hil._HilbertBase.__dict__["to_idx_array"].__wrapped__ is orig_arr
assert getattr(hil._HilbertBase.__dict__["state2idx"], "__isabstractmethod__", False)   # the abstract declaration is left alone
# host tensors still run the reference's code (hilbert.py:573-581) through the wrapper
me = types.SimpleNamespace(_idx_basis_vec=torch.tensor([2 ** n for n in range(6)]), to_idx_tensor=lambda x: torch.as_tensor(x).long())
st = torch.tensor([[1, -1, 1, -1, -1, 1], [-1, -1, -1, 1, 1, -1]], dtype=torch.int8)
assert hil._HilbertRestricted.state2idx(me, st).reshape(-1).tolist() == [1 + 4 + 32, 8 + 16]
import src.optimizer.energy as E
assert E.OptimizerBase._SGD_step is naqs_b200.energy.sgd_step
import src.naqs.wavefunction as wf
assert hasattr(wf.NAQSComplex_NADE_orbitals.sample, "__wrapped__")
print("ok")
''' % ROOT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr


def test_reference_backend_switch_leaves_reference_alone(monkeypatch):
    monkeypatch.setenv("NAQS_ELOC_BACKEND", "reference")
    assert naqs_b200.install("/nonexistent") is False
