"""Host-side output buffers for the C-ABI calls.

Large results (centrality vectors, top-k tables) are copied device->host by the library with cudaMemcpyAsync
into whatever pointer the caller passes.  Into fresh pageable memory that copy is bounded by page faults and
the driver's bounce buffers (a few GB/s); into page-locked memory it is one DMA at PCIe speed.  So the host
mirror hands the library page-locked buffers for anything big.  torch's caching host allocator supplies them
(plumbing only: freed blocks are recycled, so a steady caller pays cudaHostAlloc once); the arrays returned to
the user are plain numpy views that keep the block alive.  Small outputs, or a process without CUDA, get numpy
memory."""
import os

import numpy as np

_PIN_MIN_BYTES = 1 << 20


def host_out(shape, dtype, pinned=True):
    dtype = np.dtype(dtype)
    shape = (shape,) if np.isscalar(shape) else tuple(int(s) for s in shape)
    nbytes = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
    if pinned and nbytes >= _PIN_MIN_BYTES and not os.environ.get("SB200_PAGEABLE_OUT"):
        try:
            import torch
            if torch.cuda.is_available():
                t = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
                return t.numpy().view(dtype).reshape(shape)
        except Exception:  # no pinned memory left, or torch without CUDA: fall through to pageable memory
            pass
    return np.zeros(shape, dtype)
