"""Observation operator (`dabench.obsop.ObsOp`, dabench/obsop/_obsop.py:15-58).

The reference's default `h` indexes a state vector by the observation vector's
`location_indices` (`_index_state_vec`, :43-58).  Here that index operation is the CUDA gather
kernel (K2, `dab_obs_gather_f64`): bit-exact, for a whole ensemble at once.
"""
import numpy as np

__all__ = ['ObsOp']


class ObsOp:
    def __init__(self, h=None, H=None, location_indices=None):
        self.H = H
        self.location_indices = location_indices
        self.h = self._index_state_vec if h is None else h

    def _index_state_vec(self, state_vec, obs_vec=None):
        idx = self.location_indices if obs_vec is None else obs_vec.location_indices
        vals = np.asarray(state_vec.values if hasattr(state_vec, 'values') else state_vec, dtype=np.float64)
        return self.gather(np.atleast_2d(vals), idx).reshape(vals.shape[:-1] + (len(np.ravel(idx)),))

    @staticmethod
    def gather(X, idx, mask=None):
        """Yb[j, r] = X[j, idx[r]] on the GPU (rows with mask == 0 are zeroed)."""
        import torch
        from .._runtime import handle
        X = np.ascontiguousarray(X, dtype=np.float64)
        idx = np.ascontiguousarray(np.ravel(idx), dtype=np.int64)
        k, N = X.shape
        n_y = idx.shape[0]
        if n_y == 0:
            return np.empty((k, 0))
        if idx.min() < 0 or idx.max() >= N:
            raise IndexError('location index out of range')
        Xd, idxd = torch.from_numpy(X).cuda(), torch.from_numpy(idx).cuda()
        md = None if mask is None else torch.from_numpy(np.ascontiguousarray(mask, dtype=np.uint8)).cuda()
        Yb = torch.empty((k, n_y), dtype=torch.float64, device='cuda')
        handle().obs_gather(Xd, idxd, md, Yb, N, k, n_y, torch.cuda.current_stream().cuda_stream)
        return Yb.cpu().numpy()
