"""Device-side view of a TensorVMSplit: channel-last shadows of the VM factors, the packed alpha
mask and the POD structs of include/tensoir_b200.h.

PyTorch keeps owning the parameters in the reference NCHW layout (state_dict compatible); the
kernels read shadows that are rebuilt whenever the parameters may have changed:
  * a parameter's (data_ptr, version, shape) differs — in-place edits, shrink / upsample_volume_grid rebinding
    (tensoRF_rotated_lights.py:226-288), foreach / single-tensor optimizers;
  * any torch optimizer has stepped since the last rebuild (global post-step hook).  Fused optimizers
    (``Adam(fused=True)``, the one the CUDA-graph step uses) update parameters WITHOUT bumping the version
    counter, so the version alone is not enough;
  * the caller forces it (``refresh(model, force=True)``): every training forward does, so a captured CUDA graph
    contains the repack at the start of each replay and never depends on host-side bookkeeping.
Updates made outside torch.optim and without version bumps must be announced with :func:`mark_parameters_updated`.
"""
from __future__ import annotations

import torch

from . import _lib


def _key(ts):
    return tuple((t.data_ptr(), t._version, tuple(t.shape)) for t in ts)


# Bumped whenever parameters may have been updated behind the version counter's back (see module docstring).
_param_epoch = 0


def mark_parameters_updated(*_unused) -> None:
    """Tell every DeviceField that parameters may have changed (next ``refresh`` rebuilds the shadows)."""
    global _param_epoch
    _param_epoch += 1


def _install_optimizer_hook():
    from torch.optim.optimizer import register_optimizer_step_post_hook
    register_optimizer_step_post_hook(lambda opt, args, kwargs: mark_parameters_updated())


_install_optimizer_hook()


class DeviceField:
    def __init__(self):
        self._vm_key = None
        self._mask_key = None
        self.dplane = self.dline = self.aplane = self.aline = None
        self.amask = self.acell = None
        self.struct = _lib.TirField()
        self._keep = []

    # ---- VM factors -------------------------------------------------------------------
    def _pack(self, lib, p):
        _, Cc, H, W = p.shape
        if Cc > 1 and p.is_contiguous(memory_format=torch.channels_last):
            return p.detach()[0].permute(1, 2, 0)       # [H,W,C] view of the parameter's own storage: nothing to copy
        src = p.detach().contiguous()
        out = torch.empty((H, W, Cc), device=p.device, dtype=torch.float32)
        _lib.check(lib.tir_pack_channels_last(_lib.dptr(src), _lib.dptr(out), Cc, H, W, _lib.stream_ptr()), "pack")
        return out

    def refresh(self, model, force: bool = False) -> "_lib.TirField":
        """Bring shadows + struct up to date with ``model`` (a TensorVMSplit-like module)."""
        lib = _lib.load()
        params = list(model.density_plane) + list(model.density_line) + list(model.app_plane) + list(model.app_line)
        key = (_key(params), _param_epoch)
        s = self.struct
        if force or key != self._vm_key:
            self.dplane = [self._pack(lib, p) for p in model.density_plane]
            self.dline = [self._pack(lib, p) for p in model.density_line]
            self.aplane = [self._pack(lib, p) for p in model.app_plane]
            self.aline = [self._pack(lib, p) for p in model.app_line]
            for k in range(3):
                s.dplane[k] = self.dplane[k].data_ptr()
                s.dline[k] = self.dline[k].data_ptr()
                s.aplane[k] = self.aplane[k].data_ptr()
                s.aline[k] = self.aline[k].data_ptr()
            dcs = {p.shape[1] for p in model.density_plane}
            acs = {p.shape[1] for p in model.app_plane}
            if len(dcs) != 1 or len(acs) != 1:
                raise _lib.TirError("kernels need equal channel counts across the three orientations")
            s.dC, s.aC = dcs.pop(), acs.pop()
            self._vm_key = key
        hg = model._host_geom
        gs, lo, hi, inv = hg["grid"], hg["lo"], hg["hi"], hg["inv"]
        for i in range(3):
            s.grid[i] = gs[i]
            s.aabb_lo[i], s.aabb_hi[i], s.inv_aabb[i] = lo[i], hi[i], inv[i]
        am = getattr(model, "alphaMask", None)
        if am is None:
            s.amask = None
            s.acell = None
            self._mask_key = None
            for i in range(3):
                s.occ_lo[i], s.occ_hi[i] = 1.0, -1.0
        else:
            vol = am.alpha_volume
            mkey = _key([vol])
            if mkey != self._mask_key:
                Z, Y, X = vol.shape[-3:]
                v = vol.detach().reshape(Z, Y, X).contiguous().float()
                self.amask = torch.empty((Z, Y, X), device=v.device, dtype=torch.uint8)
                self.acell = torch.empty((Z, Y, X), device=v.device, dtype=torch.uint8)
                _lib.check(lib.tir_pack_alpha_mask(_lib.dptr(v), _lib.dptr(self.amask, torch.uint8),
                                                   _lib.dptr(self.acell, torch.uint8), X, Y, Z, _lib.stream_ptr()),
                           "pack_alpha_mask")
                self._mask_key = mkey
                # bounding box of the occupied cells (one host read per alpha-mask update, never per step)
                occ = self.acell.bool()
                self._occ = None
                if bool(occ.any()):
                    ext = []
                    for axis, n in ((2, X), (1, Y), (0, Z)):                       # x, y, z
                        hit = occ.any(dim=tuple(d for d in range(3) if d != axis)).nonzero().reshape(-1)
                        ext.append((int(hit[0]), int(hit[-1]) + 1, n))            # cells [lo, hi) along the axis
                    alo = am.aabb[0].tolist()
                    asz = (am.aabb[1] - am.aabb[0]).tolist()
                    band = 0.02                                                    # guard band in cells
                    self._occ = ([alo[i] + asz[i] * max(ext[i][0] - band, -1.0) / (ext[i][2] - 1) for i in range(3)],
                                 [alo[i] + asz[i] * min(ext[i][1] + band, ext[i][2]) / (ext[i][2] - 1) for i in range(3)])
            Z, Y, X = vol.shape[-3:]
            s.amask, s.acell = self.amask.data_ptr(), self.acell.data_ptr()
            s.agrid[0], s.agrid[1], s.agrid[2] = X, Y, Z
            alo, ainv = am._host_geom["lo"], am._host_geom["inv"]
            for i in range(3):
                s.a_lo[i], s.a_inv[i] = alo[i], ainv[i]
            occ = getattr(self, "_occ", None)
            for i in range(3):
                s.occ_lo[i], s.occ_hi[i] = (occ[0][i], occ[1][i]) if occ is not None else (1.0, -1.0)
        s.density_shift = float(model.density_shift)
        s.distance_scale = float(model.distance_scale)
        s.weight_thres = float(model.rayMarch_weight_thres)
        s.softplus = 1 if model.fea2denseAct == "softplus" else 0
        return s


def mlp_struct(model, head: str, keep: list, light: str = "index") -> "_lib.TirMlp":
    """TirMlp for one head of ``model``: 'renderModule' | 'renderModule_brdf' | 'renderModule_normal'.
    light: 'index' (light_line[light_idx]), 'mean' (mean over lights, compute_intrinfeature) or 'none'."""
    mod = getattr(model, head)
    m = _lib.TirMlp()
    ts = [mod.mlp[0].weight, mod.mlp[0].bias, mod.mlp[2].weight, mod.mlp[2].bias, mod.mlp[4].weight,
          mod.mlp[4].bias, model.basis_mat.weight]
    ts = [t.detach().contiguous() for t in ts]
    keep.extend(ts)
    m.w0, m.b0, m.w1, m.b1, m.w2, m.b2, m.basis = [t.data_ptr() for t in ts]
    ll = getattr(model, "light_line", None)
    if ll is None or light == "none":
        m.light_line, m.n_lights = None, 0
    elif light == "mean":
        # torch.mean(self.light_line(arange(light_num)), dim=0)  (tensoRF_rotated_lights.py:160-161)
        idx = torch.arange(model.light_num, device=ll.weight.device, dtype=torch.int32)
        row = torch.mean(ll(idx), dim=0).detach().contiguous().view(1, -1)
        keep.append(row)
        m.light_line, m.n_lights = row.data_ptr(), 1
    else:
        w = ll.weight.detach().contiguous()
        keep.append(w)
        m.light_line, m.n_lights = w.data_ptr(), w.shape[0]
    m.feat_dim = int(model.app_dim)
    m.hidden = int(mod.mlp[0].weight.shape[0])
    m.out_dim = int(mod.mlp[4].weight.shape[0])
    m.pe_feat = int(mod.feape)
    m.pe_x = int(getattr(mod, "viewpe", getattr(mod, "pospe", 0)))
    return m
