"""``torch.nn.functional.unfold`` (im2col) for any number of spatial dimensions.

Contract: ``einconv/functionals/unfold.py:10-56`` — note the argument order
``kernel_size, dilation, padding, stride`` (like ``F.unfold``).  The reference's einsum (``:56``)
is replaced by the bit-exact gather kernel.
"""

from typing import Tuple, Union

from torch import Tensor

from einconv_b200 import evaluator


def unfoldNd(
    x: Tensor,
    kernel_size: Union[int, Tuple[int, ...]],
    dilation: Union[int, Tuple[int, ...]] = 1,
    padding: Union[int, Tuple[int, ...], str] = 0,
    stride: Union[int, Tuple[int, ...]] = 1,
    simplify: bool = True,
) -> Tensor:
    """Extract sliding local blocks of ``x`` ``[batch_size, in_channels, *input_sizes]``.

    Returns ``[batch_size, in_channels * tot_kernel_size, tot_output_size]`` with index structure
    ``n (c_in k1 k2 ...) (o1 o2 ...)``.  ``simplify`` is accepted for API compatibility: the
    gather is one kernel for every geometry, so both settings run the same code.
    """
    del simplify
    return evaluator.unfold_autograd(
        x, kernel_size, dilation=dilation, padding=padding, stride=stride
    )


_HOST_STREAMS = {}


def unfoldNd_host(
    x_host: Tensor,
    kernel_size: Union[int, Tuple[int, ...]],
    dilation: Union[int, Tuple[int, ...]] = 1,
    padding: Union[int, Tuple[int, ...], str] = 0,
    stride: Union[int, Tuple[int, ...]] = 1,
    out: Tensor = None,
    device=None,
    chunk: int = 0,
) -> Tensor:
    """``unfoldNd`` for HOST tensors: host -> GPU -> host, pipelined over the batch.

    The unfolded tensor is ``prod(kernel_size)`` times larger than the input, so a host round trip is
    bound by the device-to-host copy.  The batch is cut into chunks that alternate between two CUDA
    streams: the upload and the gather of chunk ``i + 1`` overlap the download of chunk ``i`` (PCIe
    is full duplex and the GPU has separate copy engines per direction).  ``x_host`` and ``out``
    should be pinned (``Tensor.pin_memory()``) or the copies fall back to staged, synchronous ones.
    Returns ``out`` (allocated pinned if not given) after all work has completed.
    """
    import torch

    if x_host.is_cuda:
        raise ValueError("unfoldNd_host expects a host tensor; use unfoldNd for CUDA tensors")
    if not torch.cuda.is_available():
        raise RuntimeError("einconv_b200 has no CPU fallback: unfoldNd_host needs a CUDA device")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    B = x_host.shape[0]
    kw = dict(kernel_size=kernel_size, dilation=dilation, padding=padding, stride=stride)
    if chunk <= 0:
        # measured on C2 (scripts/time_e2e.py): 16 / 8 / 4 / 2 / 1 chunks -> 4.56 / 4.68 / 4.81 / 5.15 / 5.94 ms
        # against 3.95 ms for the bare download of the result
        chunk = max(1, B // 16)
    # the two side streams are kept per device: the caching allocator pools memory per stream, so fresh
    # streams on every call would mean fresh cudaMallocs on every call
    streams = _HOST_STREAMS.get(dev)
    if streams is None:
        streams = _HOST_STREAMS[dev] = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]
    ready = torch.cuda.Event()
    ready.record(torch.cuda.current_stream(dev))
    for i, b0 in enumerate(range(0, B, chunk)):
        s = streams[i % 2]
        s.wait_event(ready)
        with torch.cuda.stream(s):
            xd = x_host[b0 : b0 + chunk].to(dev, non_blocking=True)
            od = evaluator.unfold(xd, **kw)
            if out is None:
                shape = (B,) + tuple(od.shape[1:])
                out = torch.empty(shape, dtype=od.dtype).pin_memory()
            out[b0 : b0 + chunk].copy_(od, non_blocking=True)
            xd.record_stream(s)
            od.record_stream(s)
    for s in streams:
        s.synchronize()
    if B == 0 and out is None:
        od = evaluator.unfold(x_host.to(dev), **kw)
        out = od.cpu()
    return out
