"""torch.distributed plumbing for one-process-per-GPU jobs (torchrun): rendezvous, distribution of the
library's NCCL unique id, barriers and max-over-ranks reductions. The data path never goes through
here -- partial moments are all-reduced inside libpcf.so on the compute stream.

Replaces MPI_Init / MPI_Comm_rank / MPI_Comm_size / MPI_Finalize of the reference's _mpi programs
(reference src/mc_eur_mpi.cpp:58-66,75).
"""
from __future__ import annotations

import os
from dataclasses import dataclass


@dataclass
class Job:
    rank: int
    world: int
    local_rank: int
    backend: str | None  # None = not under torchrun (single process)


def env_job() -> tuple[int, int, int]:
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def setup(backend: str | None = None) -> Job:
    """Joins the torchrun rendezvous if there is one (WORLD_SIZE > 1); backend defaults to nccl when CUDA
    is available, gloo otherwise (CPU tests)."""
    rank, world, local = env_job()
    if world <= 1:
        return Job(0, 1, local, None)
    import torch
    import torch.distributed as td
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if backend == "nccl":
        torch.cuda.set_device(local)
    if not td.is_initialized():
        kw = {}
        if backend == "nccl":
            kw["device_id"] = torch.device("cuda", local)
        td.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return Job(rank, world, local, backend)


def share_bytes(job: Job, payload: bytes | None) -> bytes:
    """Rank 0's `payload` on every rank (used for the 128-byte NCCL unique id of libpcf.so)."""
    if job.world <= 1:
        return payload
    import torch.distributed as td
    box = [payload if job.rank == 0 else None]
    td.broadcast_object_list(box, src=0)
    return box[0]


def barrier(job: Job) -> None:
    if job.world > 1:
        import torch.distributed as td
        td.barrier()


def reduce_scalars(job: Job, values: list[float], op: str = "max") -> list[float]:
    """Element-wise max (or sum) over ranks of a short list of floats."""
    if job.world <= 1:
        return list(values)
    import torch
    import torch.distributed as td
    dev = torch.device("cuda", job.local_rank) if job.backend == "nccl" else torch.device("cpu")
    t = torch.tensor(values, dtype=torch.float64, device=dev)
    td.all_reduce(t, op=td.ReduceOp.MAX if op == "max" else td.ReduceOp.SUM)
    return t.cpu().tolist()


def teardown(job: Job) -> None:
    if job.world > 1:
        import torch.distributed as td
        if td.is_initialized():
            td.destroy_process_group()


def init_library(job: Job) -> None:
    """Brings libpcf.so up for this rank: one context on cuda:LOCAL_RANK, NCCL communicator over the job."""
    import parcompfin_b200 as pcf
    if job.world <= 1:
        pcf.init_rank(0, 1, job.local_rank, None)
        return
    nid = share_bytes(job, pcf.nccl_unique_id() if job.rank == 0 else None)
    pcf.init_rank(job.rank, job.world, job.local_rank, nid)
    # NVLink peer-memory exchange: all-gather the 64-byte CUDA IPC handles of the mailboxes; every rank must
    # succeed, otherwise the whole job stays on the NCCL all-reduce path.
    if os.environ.get("PCF_NO_PEER"):
        return
    import torch.distributed as td
    ok = 1.0
    try:
        handles = [None] * job.world
        td.all_gather_object(handles, pcf.ipc_export())
        pcf.ipc_import(handles)
    except Exception:  # noqa: BLE001 -- IPC not permitted in this environment
        ok = 0.0
    any_fail = reduce_scalars(job, [1.0 - ok], "max")[0] > 0
    if any_fail:
        pcf.peer_enable(False)
