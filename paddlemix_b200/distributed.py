"""Multi-GPU plumbing of the data-parallel path: one process per GPU (torchrun), images sharded over
ranks, ONE all_gather of the finished latents over NCCL - issued through the C ABI (`b200mix_allgather_latents`,
include/b200mix.h) on a communicator that libb200mix creates itself (`b200mix_comm_init`). torch.distributed is used only
as the launcher-side rendezvous that hands rank 0's 128-byte NCCL unique id to the other ranks.
"""
import ctypes
import glob
import os
from typing import Optional

import torch

from ._lib import check, lib

_comm: Optional[ctypes.c_void_p] = None
_world = 1
_rank = 0


def find_libnccl() -> str:
    """The libnccl.so.2 of the `nvidia-nccl` wheel torch was built against (same library torch.distributed uses), else
    the loader path."""
    env = os.environ.get("B200MIX_NCCL_LIB")
    if env:
        return env
    for base in {os.path.dirname(os.path.dirname(torch.__file__))}:
        hits = glob.glob(os.path.join(base, "nvidia", "nccl", "lib", "libnccl.so*"))
        if hits:
            return sorted(hits)[0]
    return "libnccl.so.2"


def init_comm(device: int) -> None:
    """Creates the NCCL communicator of this process (collective: every rank of the default process group must call it).
    No-op for a single process."""
    global _comm, _world, _rank
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1 or _comm is not None:
        return
    _world, _rank = dist.get_world_size(), dist.get_rank()
    check(lib.b200mix_init(int(device)), "b200mix_init")
    check(lib.b200mix_nccl_load(find_libnccl().encode()), "b200mix_nccl_load")
    uid = (ctypes.c_ubyte * 128)()
    if _rank == 0:
        check(lib.b200mix_nccl_unique_id(uid), "b200mix_nccl_unique_id")
    box = [bytes(uid)]
    dist.broadcast_object_list(box, src=0)  # the launcher's rendezvous carries the id; no tensor data moves through it
    uid = (ctypes.c_ubyte * 128).from_buffer_copy(box[0])
    comm = ctypes.c_void_p()
    check(lib.b200mix_comm_init(ctypes.byref(comm), _world, _rank, uid), "b200mix_comm_init")
    _comm = comm


def comm_ready() -> bool:
    return _comm is not None


def all_gather_latents(latents: torch.Tensor) -> torch.Tensor:
    """Finished latents of every rank -> [world * B_local, ...] (rank order) through b200mix_allgather_latents."""
    if _comm is None:
        raise RuntimeError("init_comm() has not been called on this process")
    x = latents.contiguous()
    out = torch.empty((_world * x.shape[0],) + tuple(x.shape[1:]), device=x.device, dtype=x.dtype)
    check(lib.b200mix_allgather_latents(_comm, ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(out.data_ptr()),
                                        x.numel() * x.element_size(),
                                        ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "b200mix_allgather_latents")
    return out


def destroy_comm() -> None:
    global _comm
    if _comm is not None:
        check(lib.b200mix_comm_destroy(_comm), "b200mix_comm_destroy")
        _comm = None
