"""Local-energy entry points — mirror of the hot-path methods of src/optimizer/energy.py.

`calculate_local_energy` has the signature of OptimizerBase.calculate_local_energy (energy.py:219-263)
and is what `install()` binds onto the reference's optimizer classes; `local_energy_statistics`
computes the scalars of energy.py:328,372-375 in fp64 on the device; `loss_terms` produces everything the loss of
`_SGD_step` needs (energy.py:316-329, 367-375) in one kernel, and `sgd_step` is the `_SGD_step` that uses it
(bound by `install(fused_loss=True)`).

Device-resident hand-off: when the sampler leaves its output on the GPU
(`install(device_resident=True)` sets the networks' out_device, network/base.py:29,39), `states_idx` / `psi` arrive as CUDA
tensors and nothing here touches the host: int8 rows -> naqs_state2idx -> naqs_lookup_build -> naqs_eloc -> E_loc as a
CUDA tensor.
"""
import numpy as np
import torch

from . import _lib


def np_to_torch(x):
    """src/utils/complex.py:139-140: complex numpy -> float32 [..., 2] torch (the reference truncates here)."""
    return torch.FloatTensor(np.stack([np.real(x), np.imag(x)], -1))


def _is_cuda(x):
    return torch.is_tensor(x) and x.is_cuda


def calculate_local_energy(self, states_idx, psi=None, set_unsampled_states_to_zero=True, ret_complex=False, ret_device=False):
    """Drop-in for OptimizerBase.calculate_local_energy (energy.py:219-263).

    `self` is the optimizer (needs .pauli_hamiltonian with a `.local_energy` — i.e. a PauliHamiltonianB200 —
    and, when psi is None, .wavefunction / .hilbert).  psi: complex torch [M, 2] (float32 in the reference's
    loop) or complex numpy.  Returns float32 [M, 2] torch, or complex128 numpy if ret_complex — the same
    outputs as the reference (complex.py:139-140), computed by the fused sm_100a kernel.

    CUDA inputs stay on the device: the result is then a CUDA float32 [M, 2] tensor (device-resident hand-off).
    ret_device=True (not in the reference): the untruncated CUDA float64 [M, 2] result, whatever side the inputs are on —
    what `sgd_step` / `loss_terms` consume."""
    with torch.no_grad():
        if psi is None:
            psi = self.wavefunction.psi(self.hilbert.idx2state(states_idx, use_restricted_idxs=False), ret_complex=True)
        elif torch.is_tensor(psi):
            psi = psi.detach()
        if not set_unsampled_states_to_zero:
            raise NotImplementedError()  # energy.py:250-251
        if ret_device or _is_cuda(states_idx) or _is_cuda(psi):
            eloc = self.pauli_hamiltonian.local_energy(states_idx.reshape(-1) if torch.is_tensor(states_idx) else np.asarray(states_idx).reshape(-1),
                                                       psi, ret_numpy=False)           # CUDA float64 [M, 2]
            if ret_device:
                return eloc
            if ret_complex:
                return _lib.complex_from_pairs(eloc)
            return eloc.to(torch.float32)
        idx = states_idx.detach().cpu().numpy() if torch.is_tensor(states_idx) else np.asarray(states_idx)
        local_energy = self.pauli_hamiltonian.local_energy(idx.reshape(-1), psi, ret_numpy=True)
        if not ret_complex:
            local_energy = np_to_torch(local_energy)
        return local_energy


def local_energy_statistics(table, eloc, weights=None):
    """-> dict(sum_w, mean (complex), variance of Re E) from the five all-reducible fp64 sums
    [sum w, sum w Re E, sum w Im E, sum w (Re E)^2, n] (energy.py:328,372-375)."""
    s = table.stats(eloc, weights).cpu().numpy()
    return stats_from_sums(s)


def stats_from_sums(s):
    sw = s[0]
    mean_re, mean_im = s[1] / sw, s[2] / sw
    return {"sum_w": float(sw), "mean": complex(mean_re, mean_im), "variance": float(s[3] / sw - mean_re ** 2), "n": int(round(s[4]))}


def loss_terms(table, eloc, weights=None, group=None, want=("eloc", "eloc_corr", "grad_weight")):
    """Everything the loss of `_SGD_step` needs from E_loc, computed on the device in fp64 (energy.py:316-329, 367-375):

      eloc        float32 [M, 2]  E_loc as the reference's loss sees it (complex.py:139-140)
      eloc_corr   float32 [M, 2]  E_loc - sum_i w_i E_i / sum_i w_i                      (energy.py:328)
      grad_weight float32 [M, 2]  2 w_i / sum w * conj(eloc_corr_i): exp_op == (log_psi * grad_weight).sum()  (energy.py:329)
      energy_var  float64 [3]     Re <E>, Im <E>, variance of Re E                        (energy.py:372-375)
      sums        float64 [5]     [sum w, sum w Re E, sum w Im E, sum w (Re E)^2, n], all-reduced over `group` when the
                                  batch is sharded (the only cross-rank exchange of the loss)

    eloc: CUDA float64 [M, 2] (DeviceTermTable.local_energy); weights: [M] tensor (any float dtype, any device) or None."""
    from . import distributed
    dev = table.device
    e = eloc if torch.is_tensor(eloc) else torch.from_numpy(np.ascontiguousarray(eloc))
    if e.is_complex():
        e = torch.view_as_real(e.to(torch.complex128))
    e = e.to(dev, torch.float64).contiguous()
    n = e.shape[0]
    w = None
    if weights is not None:
        w = (weights if torch.is_tensor(weights) else torch.from_numpy(np.asarray(weights))).detach().to(dev, torch.float64).reshape(-1).contiguous()
    sums = distributed.reduce_stats(table.stats(e, w), group)
    out = {k: torch.empty((n, 2), dtype=torch.float32, device=dev) for k in want}
    ev = torch.empty(3, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().naqs_loss_terms(_lib.ptr(e), _lib.ptr(w), n, _lib.ptr(sums), _lib.ptr(out.get("eloc")), _lib.ptr(out.get("eloc_corr")),
                                               _lib.ptr(out.get("grad_weight")), _lib.ptr(ev), _lib.stream_ptr(dev)), "naqs_loss_terms")
    out["energy_var"], out["sums"] = ev, sums
    return out


def sgd_step(self, states, states_idx, log_psi=None, sample_weights=None, log_psi_eval=None, regularisation_loss=None,
             n_samps=None, e_loc_clip_factor=None):
    """Drop-in for OptimizerBase._SGD_step (energy.py:273-377) with the statistics of the loss fused on the device.

    Same sequence as the reference — log amplitudes, E_loc (no gradients), << O >> = 2 Re << log_psi E_corr >>, backward,
    clip, optimizer / scheduler step, energy and variance — but steps 3 and 6 come from ONE kernel over the fp64 E_loc
    (`loss_terms`): autograd receives the detached weight vector grad_weight, so exp_op = (log_psi * grad_weight).sum() has
    exactly the gradient of energy.py:329, and (E, var) are the fp64 values instead of float32 tensor reductions.
    Works for host- and device-resident batches (E_loc stays on the device either way)."""
    ph = self.pauli_hamiltonian
    self.sampled_idxs.update(self.hilbert.to_idx_array(states_idx).squeeze())
    if log_psi is None:
        log_psi = self.wavefunction.log_psi(states)
    psi = torch.stack([log_psi.detach()[..., 0].exp() * log_psi.detach()[..., 1].cos(),
                       log_psi.detach()[..., 0].exp() * log_psi.detach()[..., 1].sin()], -1)   # cplx.exp (complex.py)
    if sample_weights is None:
        if self.reweight_samples_by_psi:
            sample_weights = log_psi.detach()[..., 0].exp().pow(2)
        else:
            raise NotImplementedError("Re-weighting by the number of samples is not yet implemented.")  # energy.py:321
    idx = states_idx.squeeze()
    with torch.no_grad():
        eloc64 = self.calculate_local_energy(idx, psi=psi, ret_device=True)
        terms = loss_terms(ph.table, eloc64, sample_weights.reshape(-1), want=("grad_weight",))
    g = terms["grad_weight"].to(log_psi.device)
    exp_op = (log_psi * g).sum()
    self.optimizer.zero_grad()
    if self.normalize_grads:
        exp_op = exp_op / (exp_op.detach()).abs()
    if regularisation_loss is not None:
        exp_op = exp_op + regularisation_loss
    exp_op.backward()
    del exp_op
    self._clip_grads()
    self.optimizer.step()
    self.optimizer.zero_grad()
    if self.scheduler is not None:
        self.scheduler.step()
    with torch.no_grad():
        if log_psi_eval is not None:  # energy.py:367-368: statistics from a second set of amplitudes
            le = log_psi_eval.detach()
            psi_eval = torch.stack([le[..., 0].exp() * le[..., 1].cos(), le[..., 0].exp() * le[..., 1].sin()], -1)
            eloc64 = self.calculate_local_energy(idx, psi=psi_eval, ret_device=True)
            terms = loss_terms(ph.table, eloc64, sample_weights.reshape(-1), want=())
        ev = terms["energy_var"].cpu().numpy()
    return float(ev[0]), float(ev[2])
