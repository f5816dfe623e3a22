"""TensorVMSplit — host-side mirror of models/tensoRF_rotated_lights.py,
models/tensoRF_general_multi_lights.py (pass ``light_name_list=``) and, as ``TensorVMSplitInit``,
models/tensoRF_init.py:139-305.  Parameters keep the reference layout and state_dict keys:
density_plane.{k} [1,C,G[m1],G[m0]], density_line.{k} [1,C,G[v],1], app_plane/app_line, basis_mat.weight,
light_line.weight, renderModule*.mlp.{0,2,4}.{weight,bias}, lgtSGs.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import tv, vm_autograd
from .tensorbase import TensorBase, AlphaGridMask, raw2alpha  # noqa: F401  (re-exported like the reference)


def _channel_last_param(values, device=None):
    """nn.Parameter with the reference's logical shape [1,C,H,W] whose STORAGE is channel-last ([H][W][C]): the layout
    the kernels index, so they read the parameter itself (device_field.py) and their channel-last gradient buffers are
    accepted by autograd without a copy.  Same random values / state_dict contents as the reference layout."""
    t = values if device is None else values.to(device)
    return torch.nn.Parameter(t.contiguous(memory_format=torch.channels_last))


class TensorVMSplit(TensorBase):
    def __init__(self, aabb, gridSize, device, **kargs):
        super().__init__(aabb, gridSize, device, **kargs)

    # ---- parameters (tensoRF_rotated_lights.py:11-29) ------------------------------------
    def init_svd_volume(self, res, device):
        self.density_plane, self.density_line = self.init_one_svd(self.density_n_comp, self.gridSize, 0.1, device)
        self.app_plane, self.app_line = self.init_one_svd(self.app_n_comp, self.gridSize, 0.1, device)
        self.basis_mat = torch.nn.Linear(sum(self.app_n_comp), self.app_dim, bias=False).to(device)
        self.light_line = torch.nn.Embedding(self.light_num, sum(self.app_n_comp)).to(device)

    def init_one_svd(self, n_component, gridSize, scale, device):
        plane_coef, line_coef = [], []
        for i in range(len(self.vecMode)):
            vec_id = self.vecMode[i]
            mat_id_0, mat_id_1 = self.matMode[i]
            plane_coef.append(_channel_last_param(
                scale * torch.randn((1, n_component[i], gridSize[mat_id_1], gridSize[mat_id_0])), device))
            line_coef.append(_channel_last_param(scale * torch.randn((1, n_component[i], gridSize[vec_id], 1)), device))
        return torch.nn.ParameterList(plane_coef), torch.nn.ParameterList(line_coef)

    def get_optparam_groups(self, lr_init_spatialxyz=0.02, lr_init_network=0.001):
        """tensoRF_rotated_lights.py:33-57 (general variant :45-46)."""
        grad_vars = [{'params': self.density_line, 'lr': lr_init_spatialxyz},
                     {'params': self.density_plane, 'lr': lr_init_spatialxyz},
                     {'params': self.app_line, 'lr': lr_init_spatialxyz},
                     {'params': self.app_plane, 'lr': lr_init_spatialxyz},
                     {'params': self.basis_mat.parameters(), 'lr': lr_init_network}]
        if hasattr(self, "light_line"):
            grad_vars += [{'params': self.light_line.parameters(), 'lr': 0.001}]
        if hasattr(self, "lgtSGs_list"):
            for p in self.lgtSGs_list:
                grad_vars += [{'params': p, 'lr': 0.001}]
        elif hasattr(self, "lgtSGs"):
            grad_vars += [{'params': self.lgtSGs, 'lr': 0.001}]
        grad_vars += [{'params': self.renderModule.parameters(), 'lr': lr_init_network}]
        if hasattr(self, "renderModule_brdf"):
            grad_vars += [{'params': self.renderModule_brdf.parameters(), 'lr': lr_init_network}]
        if hasattr(self, "renderModule_normal") and self.normals_kind in (
                "purely_predicted", "derived_plus_predicted", "residue_prediction"):
            grad_vars += [{'params': self.renderModule_normal.parameters(), 'lr': lr_init_network}]
        return grad_vars

    # ---- regularisers (parameter-only, same values as tensoRF_rotated_lights.py:60-92) --------------
    @staticmethod
    def _offdiag_gram_mean(line):
        """mean |<v_i, v_j>|, i != j, over the components of one line factor [1,C,D,1] (the 'Ortho' term)."""
        v = line.reshape(line.shape[1], line.shape[2])
        gram = v @ v.t()
        off = ~torch.eye(gram.shape[0], dtype=torch.bool, device=gram.device)
        return gram[off].abs().mean()

    def vectorDiffs(self, vector_comps):
        return sum(self._offdiag_gram_mean(v) for v in vector_comps)

    def vector_comp_diffs(self):
        return self.vectorDiffs(self.density_line) + self.vectorDiffs(self.app_line)

    def density_L1(self):
        return sum(p.abs().mean() + l.abs().mean() for p, l in zip(self.density_plane, self.density_line))

    def TV_loss_density(self, reg):
        """tensoRF_rotated_lights.py:80-85; one kernel launch per direction of autograd (tv.py)."""
        return tv.tv_planes(self.density_plane, reg, 1e-2, self.__dict__.get("_tir_grad_inplace", True))

    def TV_loss_app(self, reg):
        """tensoRF_rotated_lights.py:87-92."""
        return tv.tv_planes(self.app_plane, reg, 1e-2, self.__dict__.get("_tir_grad_inplace", True))

    # ---- VM gathers (kernel-backed, differentiable w.r.t. the factors) -------------------
    def compute_densityfeature(self, xyz_sampled):
        """tensoRF_rotated_lights.py:95-110."""
        return vm_autograd.density_feature(self, xyz_sampled)

    def compute_densityfeature_with_xyz_grad(self, xyz_sampled):
        """tensoRF_rotated_lights.py:113-129; the spatial gradient is available from
        vm_autograd.density_feature_and_grad (analytic, no autograd.grad round trip)."""
        return vm_autograd.density_feature_and_grad(self, xyz_sampled)[0]

    def _light_rows(self, light_idx, n):
        if light_idx is None:
            return None
        return self.light_line(light_idx.reshape(-1).to(self.light_line.weight.device))

    def _mean_light(self):
        idx = torch.arange(self.light_num, device=self.light_line.weight.device, dtype=torch.int32)
        return torch.mean(self.light_line(idx), dim=0)

    def compute_bothfeature(self, xyz_sampled, light_idx):
        """tensoRF_rotated_lights.py:132-165 -> (radiance_field_feat, intrinsic_feat)."""
        prod = vm_autograd.app_products(self, xyz_sampled)          # [n, 3*C] = plane*line
        rad = self.basis_mat(prod * self._light_rows(light_idx, prod.shape[0]))
        intr = self.basis_mat(prod * self._mean_light()[None, :])
        return rad, intr

    def compute_intrinfeature(self, xyz_sampled):
        """tensoRF_rotated_lights.py:167-195."""
        prod = vm_autograd.app_products(self, xyz_sampled)
        return self.basis_mat(prod * self._mean_light()[None, :])

    def compute_appfeature(self, xyz_sampled, light_idx):
        """tensoRF_rotated_lights.py:197-224."""
        prod = vm_autograd.app_products(self, xyz_sampled)
        return self.basis_mat(prod * self._light_rows(light_idx, prod.shape[0]))

    # ---- grid maintenance (behaviour of tensoRF_rotated_lights.py:226-288) ---------------------------------
    # Both operations REBIND the Parameters (the train loop rebuilds Adam afterwards, train_tensoIR.py:421-422);
    # DeviceField notices the new (data_ptr, version, shape) and rebuilds the kernel-side shadows.
    def _rebind(self, plist, k, tensor):
        plist[k] = _channel_last_param(tensor)

    @torch.no_grad()
    def up_sampling_VM(self, plane_coef, line_coef, res_target):
        for k, (m, v) in enumerate(zip(self.matMode, self.vecMode)):
            size_plane = (res_target[m[1]], res_target[m[0]])
            self._rebind(plane_coef, k, F.interpolate(plane_coef[k].data, size=size_plane, mode='bilinear',
                                                      align_corners=True))
            self._rebind(line_coef, k, F.interpolate(line_coef[k].data, size=(res_target[v], 1), mode='bilinear',
                                                     align_corners=True))
        return plane_coef, line_coef

    @torch.no_grad()
    def upsample_volume_grid(self, res_target):
        for planes, lines in ((self.app_plane, self.app_line), (self.density_plane, self.density_line)):
            self.up_sampling_VM(planes, lines, res_target)
        self.update_stepSize(res_target)
        print(f'upsamping to {res_target}')

    @torch.no_grad()
    def shrink(self, new_aabb):
        """Crop every factor to the voxel range covering ``new_aabb`` and snap the aabb to that range when the
        alpha-mask resolution differs from the grid's."""
        print("====> shrinking ...")
        lo = torch.round(torch.round((new_aabb[0] - self.aabb[0]) / self.units)).long()
        hi = torch.minimum(torch.round((new_aabb[1] - self.aabb[0]) / self.units).long() + 1, self.gridSize)
        for k, (m, v) in enumerate(zip(self.matMode, self.vecMode)):
            for lines in (self.density_line, self.app_line):
                self._rebind(lines, k, lines[k].data[..., lo[v]:hi[v], :])
            for planes in (self.density_plane, self.app_plane):
                self._rebind(planes, k, planes[k].data[..., lo[m[1]]:hi[m[1]], lo[m[0]]:hi[m[0]]])
        if not torch.all(self.alphaMask.gridSize == self.gridSize):
            f_lo, f_hi = lo / (self.gridSize - 1), (hi - 1) / (self.gridSize - 1)
            snapped = torch.stack(((1 - f_lo) * self.aabb[0] + f_lo * self.aabb[1],
                                   (1 - f_hi) * self.aabb[0] + f_hi * self.aabb[1]))
            print("aabb", new_aabb, "\ncorrect aabb", snapped)
            new_aabb = snapped
        self.aabb = new_aabb
        self.update_stepSize(tuple((hi - lo).tolist()))
