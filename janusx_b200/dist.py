"""SNP-range sharding of the scan across the GPUs of one box (one process per GPU, torch.distributed).

SNPs are independent given the null model (U^T, S, Xcov, y_rot, bounds, nullml) -- the reference exploits
exactly this with rayon (src/stats/reml.rs:88-99).  So the multi-GPU path is:
  1. rank 0 holds the null model (eigendecomposition done once); ONE broadcast ships it to every rank
     (NCCL over NVLink for the 4*n*n-byte U^T when the tensors live on GPUs; gloo on CPU in tests);
  2. every rank scans a contiguous SNP range [g*m/G, (g+1)*m/G) of the BED in BED order
     (byte offsets 3 + snp*ceil(n_full/4), src/stats/lmm.rs:1213-1215) -- no data-path collective;
  3. results are gathered in rank order (= BED order, filtered rows removed) on rank 0.
No reduction across GPUs exists anywhere on the path, and per-SNP results never depend on G (no warm start).
"""
from __future__ import annotations

import os
import shutil
from dataclasses import dataclass
from typing import Callable, Optional, Sequence, Tuple

import numpy as np


def shard_range(m: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous SNP range of `rank`: [rank*m/world, (rank+1)*m/world)."""
    return (rank * m) // world, ((rank + 1) * m) // world


@dataclass
class NullModel:
    s: np.ndarray        # f64[n]
    xcov: np.ndarray     # f64[n, p] rotated design
    y: np.ndarray        # f64[n] rotated phenotype
    u_t: object          # f32[n, n]: numpy array, or a torch tensor (CUDA after an NCCL broadcast)
    low: float
    high: float
    lbd_null: float
    nullml: float = float("nan")


def env_rank_world() -> Tuple[int, int, int]:
    return int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))


def init_process_group(backend: Optional[str] = None, device_index: Optional[int] = None):
    """Rendezvous from RANK/WORLD_SIZE/MASTER_ADDR/MASTER_PORT (torchrun).  backend: nccl on GPUs (one GPU per rank),
    gloo on CPU or when ranks share a GPU.  device_index: this rank's CUDA device (default LOCAL_RANK)."""
    import torch
    import torch.distributed as dist

    rank, world, local = env_rank_world()
    if world == 1 or dist.is_initialized():
        return dist if dist.is_initialized() else None
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    dev = local if device_index is None else int(device_index)
    if backend == "nccl":
        torch.cuda.set_device(dev)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{dev}"))
    else:
        if torch.cuda.is_available():
            torch.cuda.set_device(dev)
        dist.init_process_group("gloo")
    return dist


def broadcast_null_model(model: Optional[NullModel], n: int, p: int, src: int = 0, device=None) -> NullModel:
    """One broadcast of (S, Xcov, y_rot, scalars) + one of U^T (f32).  Ranks != src pass model=None."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size() == 1:
        assert model is not None
        return model
    rank = dist.get_rank()
    use_cuda = dist.get_backend() == "nccl"
    dev = device if device is not None else (torch.device(f"cuda:{torch.cuda.current_device()}") if use_cuda
                                             else torch.device("cpu"))
    small = torch.empty(n + n * p + n + 4, dtype=torch.float64, device=dev)
    if rank == src:
        flat = np.concatenate([model.s, model.xcov.reshape(-1), model.y,
                               [model.low, model.high, model.lbd_null, model.nullml]])
        small.copy_(torch.as_tensor(flat))
        ut = model.u_t if hasattr(model.u_t, "data_ptr") else torch.as_tensor(np.ascontiguousarray(model.u_t))
        ut = ut.to(dev, dtype=torch.float32).contiguous()
    else:
        ut = torch.empty((n, n), dtype=torch.float32, device=dev)
    dist.broadcast(small, src)
    dist.broadcast(ut, src)      # 4*n*n bytes, once (NVLink / NVSwitch under NCCL)
    sm = small.cpu().numpy()
    o = 0
    s = sm[o:o + n].copy(); o += n
    xcov = sm[o:o + n * p].reshape(n, p).copy(); o += n * p
    y = sm[o:o + n].copy(); o += n
    low, high, lbd, nullml = (float(v) for v in sm[o:o + 4])
    return NullModel(s=s, xcov=xcov, y=y, u_t=(ut if use_cuda else ut.numpy()), low=low, high=high, lbd_null=lbd,
                     nullml=nullml)


def gather_rows_in_order(local_rows: np.ndarray, dst: int = 0) -> Optional[np.ndarray]:
    """Ordered gather of per-rank result rows (f64[k_r, cols]) to rank `dst`: concatenation in rank order."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local_rows
    world, rank = dist.get_world_size(), dist.get_rank()
    use_cuda = dist.get_backend() == "nccl"
    dev = torch.device(f"cuda:{torch.cuda.current_device()}") if use_cuda else torch.device("cpu")
    cols = int(local_rows.shape[1])
    counts = torch.zeros(world, dtype=torch.int64, device=dev)
    counts[rank] = local_rows.shape[0]
    dist.all_reduce(counts)          # bookkeeping only: row counts, not data
    cnt = [int(c) for c in counts.cpu()]
    kmax = max(cnt) if cnt else 0
    pad = torch.zeros((kmax, cols), dtype=torch.float64, device=dev)
    if local_rows.shape[0]:
        pad[: local_rows.shape[0]] = torch.as_tensor(np.ascontiguousarray(local_rows), device=dev)
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst)
    if rank != dst:
        return None
    return np.concatenate([b[:c].cpu().numpy() for b, c in zip(bufs, cnt)], axis=0)


def scan_bed_sharded(bed_prefix: str, out_tsv: str, n_snps: int, scan_range: Callable[[int, int, str, bool], int],
                     barrier: bool = True) -> int:
    """Every rank scans its contiguous SNP range into `<out_tsv>.part<rank>` with
    scan_range(snp_begin, snp_end, part_path, write_header) -> rows written; rank 0 then concatenates the parts
    in rank order (= BED order).  Returns the total row count on every rank."""
    import torch
    import torch.distributed as dist

    rank, world = (dist.get_rank(), dist.get_world_size()) if dist.is_initialized() else (0, 1)
    b, e = shard_range(n_snps, rank, world)
    part = f"{out_tsv}.part{rank}" if world > 1 else out_tsv
    rows = int(scan_range(b, e, part, rank == 0))
    if world == 1:
        return rows
    use_cuda = dist.get_backend() == "nccl"
    dev = torch.device(f"cuda:{torch.cuda.current_device()}") if use_cuda else torch.device("cpu")
    tot = torch.tensor([rows], dtype=torch.int64, device=dev)
    dist.all_reduce(tot)             # doubles as the completion barrier ...
    total = int(tot.item())          # ... once the HOST has seen its result: an NCCL collective returns as soon as it is
                                     # enqueued, and rank 0 must not read a part file another rank is still writing
    if rank == 0:
        with open(out_tsv, "wb") as out:
            for r in range(world):
                p = f"{out_tsv}.part{r}"
                with open(p, "rb") as fh:
                    shutil.copyfileobj(fh, out, 8 << 20)
                os.remove(p)
    if barrier:
        dist.barrier()
    return total
