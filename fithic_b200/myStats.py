"""Mirror of the reference's fithic/myStats.py for the one function on the hot path."""
import numpy as np


def benjamini_hochberg_correction(p_values, num_total_tests):
    """Same contract as myStats.benjamini_hochberg_correction (reference fithic/myStats.py:24-48): list of p-values and
    the number of tests in, list of q-values in input order out.  The ranking (sort), the forward running max of
    min(1, p*T/rank) and the scatter run on the GPU (fhc_bh_qvalues); there is no CPU fallback."""
    import torch
    from . import _capi
    from ._capi import check, dptr
    lib = _capi.load()
    p = torch.as_tensor(np.asarray(p_values, dtype=np.float64)).cuda()
    n = p.numel()
    q = torch.empty(max(n, 1), dtype=torch.float64, device=p.device)[:n]
    wsb = int(lib.fhc_bh_workspace_bytes(n))
    ws = torch.empty(wsb, dtype=torch.uint8, device=p.device)
    check(lib.fhc_bh_qvalues(dptr(p), n, float(num_total_tests), 0, 0.0, dptr(q), None, None, dptr(ws), wsb,
                             torch.cuda.current_stream().cuda_stream))
    return q.cpu().tolist()
