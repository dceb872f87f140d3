"""Strip-sharded execution of one large image / volume over the ranks of a torch.distributed job.

The reference is single-GPU; this layer is new.  A 2-D image is cut into
`world` horizontal strips (a volume into z slabs), one per rank / GPU.  Scans along the other
dimensions are strip local.  For the scans along the cut dimension every rank

  stage 1   filters its strip with zero incoming carries (rf_plan_stage1) and obtains, per line
            crossing the cut, the order-r tail of every scan: `shard_tail_bytes` bytes
            (2 scans x r=3 x 8192 columns x 4 B = 196 KB for the headline Gaussian);
  exchange  the tails of every strip reach every rank: by default peer to peer over NVLink, straight from
            device memory into the other ranks' exchange windows (rf_xchg_put / rf_xchg_wait: CUDA IPC
            mappings, no collective library and no host in the data path); alternatively ONE NCCL
            all-gather, or two column-chunked all-to-alls (gloo in the CPU tests); or, for filters that forget a
            whole strip (rf_plan_shard_neighbors_suffice), the windows of the two adjacent ranks only
            (exchange="neighbor": rf_xchg_put_part / rf_xchg_wait_from);
  stage 2   resolves the carries entering its strip from the gathered tails with the whole-strip
            transition matrices (a tiny kernel, redundantly on every rank), corrects the stage-1 carries of
            the strip with them (no second carry chain) and finishes the filter (rf_plan_stage2).

No image data crosses the link.  Batches of independent images need no exchange at all.
"""
from __future__ import annotations

from typing import Sequence

import torch
import torch.distributed as dist

from .capi import Exchange, Plan, Scan


def exchange_tails(my_tails: torch.Tensor, world: int, group=None) -> torch.Tensor:
    """All-gather per-image tails: [B, E] on every rank -> [B, world, E], rank-major per image
    (the layout rf_plan_stage2 expects for one image is [world, E])."""
    if my_tails.dim() != 2:
        raise ValueError("my_tails must be [images, tail elements]")
    if world == 1:
        return my_tails.unsqueeze(1).contiguous()
    b, e = my_tails.shape
    # concatenated output form (world*B, E): accepted by both the NCCL and the gloo backend
    gathered = torch.empty((world * b, e), dtype=my_tails.dtype, device=my_tails.device)
    dist.all_gather_into_tensor(gathered, my_tails.contiguous(), group=group)
    return gathered.view(world, b, e).transpose(0, 1).contiguous()


def scatter_tail_chunks(my_tails: torch.Tensor, vectors: int, world: int, group=None) -> torch.Tensor:
    """Column-chunked exchange, step 1.  my_tails is this shard's [vectors * lines] tail array; every rank
    receives the tails of ALL shards for ITS chunk of the lines: returns [world, vectors, lines // world]
    (shard-major: the layout rf_plan_shard_resolve_lines expects)."""
    lines = my_tails.numel() // vectors
    c = lines // world
    send = my_tails.view(vectors, world, c).permute(1, 0, 2).contiguous()        # [dest rank][vectors][c]
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send, group=group)
    return recv


def gather_carry_chunks(ext_all: torch.Tensor, group=None) -> torch.Tensor:
    """Column-chunked exchange, step 2.  ext_all is [world (target shard), vectors, c]: the carries entering
    every shard for this rank's line chunk.  Returns this shard's full [vectors * lines] carry array."""
    world, vectors, c = ext_all.shape
    back = torch.empty_like(ext_all)
    dist.all_to_all_single(back, ext_all.contiguous(), group=group)              # back[q] = my carries, chunk q
    return back.permute(1, 0, 2).contiguous().view(vectors * world * c)


def strip_bounds(extent: int, world: int, rank: int, multiple: int = 1) -> tuple[int, int]:
    """[lo, hi) of rank's strip along the cut dimension: equal strips, `multiple` keeps tile alignment."""
    if extent % (world * multiple) != 0:
        raise ValueError(f"extent {extent} is not divisible into {world} strips that are multiples of {multiple}")
    n = extent // world
    return rank * n, (rank + 1) * n


class ShardedFilter:
    """One rank's share of a strip-sharded filter over a batch of `batch` images.

    extents    full extents, dimension 0 (contiguous) first
    shard_dim  dimension that is cut (must carry at least one scan; not 0)
    """

    def __init__(self, extents: Sequence[int], dtype, scans: Sequence[Scan], border: str, *, rank: int, world: int,
                 shard_dim: int | None = None, batch: int = 1, engine: str = "auto", group=None, stacked: bool = False,
                 overlap: int = 1, exchange: str = "auto", graph: bool = False):
        """stacked=True: the `batch` images are one dense stack [batch][...] and are filtered as ONE filter with
        an extra outermost dimension that carries no scans (allowed by the reference: lib/split.cpp:1888-1898,
        it is how apps/audio batches channels): one launch sequence and one tail exchange per stack.
        overlap=g > 1 (stacked only): the stack is split into g sub-stacks, each with its own plan (carry
        workspace) and CUDA stream, so the latency-bound carry stage of one sub-stack runs beside the
        bandwidth-bound tile kernels of another.
        exchange: "p2p" (every rank writes its tails straight into every other rank's exchange window over
        NVLink -- rf_xchg_put / rf_xchg_wait, no NCCL call per step), "allgather" (one NCCL all-gather),
        "alltoall" (column-chunked: every rank resolves 1/world of the lines for all shards; 2 small all-to-alls
        instead of one all-gather whose volume grows with world) or "auto" (the collectives: all-gather, all-to-all
        from 4 ranks on when the all-gather would deliver 8 MB or more per rank -- measured on 8 B200s the
        peer-to-peer windows tie with the NCCL all-gather, and only the collectives can be captured in a CUDA graph).
        graph=True (stacked only): the whole step -- stage 1, the exchange, stage 2 -- is captured once per
        (src, dst) pair in a CUDA graph and replayed: a sharded step is a dozen short launches, and from 4 ranks on
        the host cannot issue them as fast as the GPUs finish them."""
        self.rank, self.world, self.group, self.batch = rank, world, group, batch
        self.stacked = stacked and batch > 1
        self.groups = overlap if (self.stacked and overlap > 1 and batch % overlap == 0) else 1
        extents = list(int(e) for e in extents)
        self.shard_dim = len(extents) - 1 if shard_dim is None else shard_dim
        lo, hi = strip_bounds(extents[self.shard_dim], world, rank)
        self.local_extents = list(extents)
        self.local_extents[self.shard_dim] = hi - lo
        self.rows = (lo, hi)
        kw = dict(engine=engine)
        if world > 1:
            kw.update(shard_dim=self.shard_dim, open_lo=rank > 0, open_hi=rank < world - 1)
        if self.stacked:
            self.sub = batch // self.groups
            self.plans = [Plan(self.local_extents + [self.sub], dtype, scans, border, **kw) for _ in range(self.groups)]
            self.streams = [torch.cuda.Stream() for _ in range(self.groups)] if self.groups > 1 else []
        else:
            # one plan per image in flight: a plan owns the carry workspace of its image
            self.plans = [Plan(self.local_extents, dtype, scans, border, **kw) for _ in range(batch if world > 1 else 1)]
        self.shard_causal = [bool(sc.causal) for sc in scans if int(sc.dim) == self.shard_dim]
        self.tail_elems = self.plans[0].shard_tail_bytes // 4 if world > 1 else 0
        self._tails = None
        self.vectors = self.plans[0].shard_vectors if world > 1 else 0
        lines = self.tail_elems // self.vectors if self.vectors else 0
        chunked_ok = world > 1 and self.vectors > 0 and lines % world == 0
        if exchange == "alltoall" and not chunked_ok:
            raise ValueError("exchange='alltoall' needs the cut lines to divide evenly among the ranks")
        # measured on 4 and 8 B200s (profiles/): the two all-to-alls win once the all-gather would deliver >= ~8 MB
        # per rank; below that its single collective is faster
        big = self.tail_elems * 4 * world >= (8 << 20)
        # neighbour exchange: peer-to-peer windows, but a rank only sends the causal scans' tails to rank + 1 and the
        # anticausal ones to rank - 1 (and waits for those two): valid when every rank's plan reports that the adjacent
        # strips decide its carries (short memory over a whole strip)
        self.neighbor = False
        if world > 1 and exchange == "neighbor":
            if not torch.cuda.is_available():
                raise ValueError("exchange='neighbor' needs CUDA devices")
            ok = torch.tensor([1 if all(p.shard_neighbors_suffice for p in self.plans) else 0], device="cuda", dtype=torch.int32)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            if int(ok.item()) != 1:
                raise ValueError("exchange='neighbor': the filter remembers more than one strip (rf_plan_shard_neighbors_suffice is 0)")
            self.neighbor = True
        self.p2p = world > 1 and exchange in ("p2p", "neighbor") and torch.cuda.is_available()
        self.use_graph = bool(graph) and self.stacked and not self.p2p and torch.cuda.is_available()
        self._graphs = {}
        self.chunked = (not self.p2p) and chunked_ok and (exchange == "alltoall" or (exchange == "auto" and world >= 4 and big))
        self.windows = []
        if self.p2p:
            # one exchange window per plan in flight; CUDA IPC handles travel once, here
            for _ in self.plans:
                x = Exchange(self.tail_elems * 4, world, rank)
                x.connect(group)
                self.windows.append(x)

    def _finish(self, plan, src, dst, tails):
        """Exchange the strip tails and finish the filter (stage 2) for one plan."""
        if self.neighbor:
            x = self.windows[self.plans.index(plan)]
            me, lo, hi = 1 << self.rank, self.rank - 1, self.rank + 1
            nbytes = self.tail_elems * 4
            per_scan = nbytes // len(self.shard_causal)               # [scan][order][lines]: one contiguous block per scan
            x.put_part(me, tails, 0, nbytes, False)                   # own slot: the resolver reads this strip's tails too
            mask = me
            for s, causal in enumerate(self.shard_causal):
                to = hi if causal else lo
                if 0 <= to < self.world:
                    x.put_part(1 << to, tails, s * per_scan, per_scan, False)
            for nb in (lo, hi):
                if 0 <= nb < self.world:
                    mask |= 1 << nb
            x.put_part(mask, None, 0, 0, True)                         # arrival words: own window and both neighbours'
            plan.stage2(src, dst, x.wait_from(mask), self.world, self.rank)
        elif self.p2p:
            x = self.windows[self.plans.index(plan)]
            x.put(tails)
            plan.stage2(src, dst, x.wait(), self.world, self.rank)
        elif self.chunked:
            recv = scatter_tail_chunks(tails, self.vectors, self.world, self.group)
            ext_all = torch.empty_like(recv)
            plan.shard_resolve_lines(recv, self.world, recv.shape[2], ext_all)
            plan.stage2_ext(src, dst, gather_carry_chunks(ext_all, self.group))
        else:
            gathered = exchange_tails(tails.view(1, -1), self.world, self.group)
            plan.stage2(src, dst, gathered[0], self.world, self.rank)

    def run_stacked(self, src: torch.Tensor, dst: torch.Tensor):
        """Filter a dense stack [batch][local extents...] (stacked=True)."""
        if not self.stacked:
            raise ValueError("run_stacked needs stacked=True")
        if not self.use_graph:
            return self._run_stacked(src, dst)
        key = (src.data_ptr(), dst.data_ptr())
        g = self._graphs.get(key)
        if g is None:
            self._run_stacked(src, dst)              # lazy initialisations (function attributes, NCCL) stay outside the capture
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._run_stacked(src, dst)
            self._graphs[key] = g
        g.replay()

    def _run_stacked(self, src: torch.Tensor, dst: torch.Tensor):
        G, sub = self.groups, self.sub
        if self.world > 1 and self._tails is None:
            dt = torch.float32 if src.dtype == torch.float32 else torch.int32
            self._tails = torch.empty((G, self.tail_elems), device=src.device, dtype=dt)
        if G == 1:
            plan = self.plans[0]
            if self.world == 1:
                plan.execute(src, dst)
                return
            plan.stage1(src, dst, self._tails[0])
            self._finish(plan, src, dst, self._tails[0])
            return
        # sub-stacks on their own streams; the caller's stream is joined at both ends
        cur = torch.cuda.current_stream()
        parts = [(src[g * sub:(g + 1) * sub], dst[g * sub:(g + 1) * sub]) for g in range(G)]
        for st in self.streams:
            st.wait_stream(cur)
        if self.world == 1:
            for g, (a, b) in enumerate(parts):
                with torch.cuda.stream(self.streams[g]):
                    self.plans[g].execute(a, b)
        else:
            for g, (a, b) in enumerate(parts):
                with torch.cuda.stream(self.streams[g]):
                    self.plans[g].stage1(a, b, self._tails[g])
            for g, (a, b) in enumerate(parts):       # collectives are issued in the same order on every rank
                with torch.cuda.stream(self.streams[g]):
                    self._finish(self.plans[g], a, b, self._tails[g])
        for st in self.streams:
            cur.wait_stream(st)

    def run(self, srcs: Sequence[torch.Tensor], dsts: Sequence[torch.Tensor]):
        """Filter `batch` strips (device tensors of the local extents)."""
        if self.world == 1:
            for s, d in zip(srcs, dsts):
                self.plans[0].execute(s, d)
            return
        if self._tails is None:
            dt = torch.float32 if srcs[0].dtype == torch.float32 else torch.int32
            self._tails = torch.empty((self.batch, self.tail_elems), device=srcs[0].device, dtype=dt)
        for i, (s, d) in enumerate(zip(srcs, dsts)):
            self.plans[i].stage1(s, d, self._tails[i])
        if self.p2p:
            for i, (s, d) in enumerate(zip(srcs, dsts)):
                self._finish(self.plans[i], s, d, self._tails[i])
            return
        gathered = exchange_tails(self._tails, self.world, self.group)
        for i, (s, d) in enumerate(zip(srcs, dsts)):
            self.plans[i].stage2(s, d, gathered[i], self.world, self.rank)

    @property
    def launches_per_image(self) -> int:
        # strip resolve + carry correction; p2p: + the two one-warp arrival kernels
        n = self.plans[0].num_launches + ((2 + (2 if self.p2p else 0)) if self.world > 1 else 0)
        return n * self.groups / self.batch if self.stacked else n

    def close(self):
        self._graphs = {}
        if self.windows:
            torch.cuda.synchronize()
            if dist.is_initialized():
                dist.barrier(group=self.group)       # nobody unmaps a window a peer may still write to
            for x in self.windows:
                x.close()
            self.windows = []
        for p in self.plans:
            p.close()
