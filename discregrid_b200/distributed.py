"""Multi-GPU sharding of the addFunction node loop: one process per GPU, torch.distributed for the plumbing.

Nodes are independent (cubic_lagrange_discrete_grid.cpp:806-817 is an embarrassingly parallel loop; OpenMP already
splits it statically), so the node index space [0, n) is cut into `rows * world` equal chunks dealt round-robin:
rank r owns chunks r, r + world, r + 2*world, ...  Round-robin (rather than one contiguous slab per rank) evens out
the spatially varying cost of the nearest-triangle query (far-from-surface nodes prune worse).  Two rows per rank are the
measured optimum (every extra launch adds a ~2 ms tail of long-running warps and partially masked bricks).  Row j of the deal
is gathered by ONE all-gather whose output is the contiguous slice [j*world*chunk, (j+1)*world*chunk) of the full
coefficient array -- so the gathered array is already in the reference's node order and no permutation pass is
needed.  The only exchange step on the path is this all-gather of 8-byte coefficients (NCCL over NVLink on the GPU
box, gloo in the CPU tests).
"""
from dataclasses import dataclass


@dataclass(frozen=True)
class NodeSharding:
    n: int            # number of nodes
    world: int
    rows: int
    chunk: int        # nodes per chunk (last chunks may be short / empty)

    @property
    def padded(self):
        return self.rows * self.world * self.chunk

    def chunks_of(self, rank):
        """[(row, l_begin, l_end)] owned by `rank` (ranges clipped to n; may be empty)."""
        out = []
        for j in range(self.rows):
            b = (j * self.world + rank) * self.chunk
            out.append((j, min(b, self.n), min(b + self.chunk, self.n)))
        return out



def _stream_ptr(t):
    """the current CUDA stream's handle for a device tensor; the default stream (0) for a host tensor -- host tensors only occur when the
    multi-GPU logic is rehearsed on the CPU against the emulated library (tests/emu, gloo)"""
    import ctypes as C
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream) if t.is_cuda else C.c_void_p(0)


def _side_streams(t, n):
    """n side streams on t's device; none for a host tensor (its launches have completed when they return)"""
    import torch
    return [torch.cuda.Stream() for _ in range(n)] if (t.is_cuda and n > 0) else []

def make_sharding(n, world, rows=None, align=1024):
    """Equal chunks of a size that is a multiple of `align` nodes (keeps every chunk 8 KiB-aligned in the fp64 array)."""
    if world == 1:
        rows = 1
    elif rows is None:
        rows = 2                           # measured at 8 ranks, 128^3: rows 1/2/4/8 -> slowest rank 10.6/9.7/10.4/11.0 ms (ideal 6.8):
                                           # more rows = better balance but more launch tails and masked brick planes
    per = -(-n // (rows * world))
    chunk = -(-per // align) * align
    return NodeSharding(n, world, rows, chunk)


def allgather_rows(full, sharding, group=None):
    """`full`: 1-D tensor of `sharding.padded` elements in which this rank has filled its own chunks.
    After the call every rank holds all chunks.  One collective per row; each lands in its final position."""
    import torch.distributed as dist
    if sharding.world == 1:
        return
    rank = dist.get_rank(group)
    w, c = sharding.world, sharding.chunk
    for j in range(sharding.rows):
        row = full[j * w * c:(j + 1) * w * c]
        dist.all_gather_into_tensor(row, row[rank * c:(rank + 1) * c].clone() if row.device.type == "cpu" else row[rank * c:(rank + 1) * c], group=group)


class ShardedSdfSampler:
    """addFunction's node loop for one rank of a one-process-per-GPU job: this rank's chunks are launched round-robin on a
    few side streams (a K1 launch ends with a ~2 ms tail of long-running warps; concurrent launches hide each other's
    tails), joined, then the rows are all-gathered.  Plumbing only: the kernels are the C-ABI's dg_sample_sdf_device."""

    def __init__(self, md, desc, sharding, rank, n_streams=4):
        import torch
        self.md, self.desc, self.sh, self.rank = md, desc, sharding, rank
        self.chunks = [(j, b, e) for (j, b, e) in sharding.chunks_of(rank) if e > b]
        self.n_streams = min(n_streams, max(1, len(self.chunks))) if len(self.chunks) > 1 else 0
        self.streams = None                                # created at the first launch, on the output tensor's device

    def launch(self, full, sign=1.0):
        """enqueue this rank's kernels; `full` = fp64 CUDA tensor of sharding.padded elements; returns #launches"""
        import ctypes as C
        import torch
        from . import _capi as capi
        if self.streams is None:
            self.streams = _side_streams(full, self.n_streams)
        if not self.streams:
            for (_j, b, e) in self.chunks:
                capi.check(capi.lib.dg_sample_sdf_device(self.md.handle, C.byref(self.desc), sign, b, e,
                                                         C.c_void_p(full.data_ptr() + 8 * b), _stream_ptr(full)))
            return len(self.chunks)
        cur = torch.cuda.current_stream()
        for st in self.streams:
            st.wait_stream(cur)
        for k, (_j, b, e) in enumerate(self.chunks):
            st = self.streams[k % len(self.streams)]
            capi.check(capi.lib.dg_sample_sdf_device(self.md.handle, C.byref(self.desc), sign, b, e,
                                                     C.c_void_p(full.data_ptr() + 8 * b), C.c_void_p(st.cuda_stream)))
        for st in self.streams:
            cur.wait_stream(st)
        return len(self.chunks)

    def step(self, full, sign=1.0, group=None):
        n = self.launch(full, sign)
        allgather_rows(full, self.sh, group)
        return n


class ShardedDensityMap:
    """GenerateDensityMap's node loop (K3, dg_density_map_device) for one rank of a one-process-per-GPU job: the SDF field is
    replicated (every rank holds the gathered coefficients anyway), the nodes are dealt in `rows` round-robin node-id chunks per rank
    -- many small chunks, because only the nodes in the surface band take the 4096-point branch -- and the density coefficients are
    exchanged exactly like the SDF's: one in-place all-gather per row ("density map = same all-gather as K1")."""

    def __init__(self, field_handle, n_nodes, rank, world, rows=8):
        self.fh, self.rank, self.world = field_handle, rank, world
        self.sh = make_sharding(n_nodes, world, rows)
        self.chunks = [(j, b, e) for (j, b, e) in self.sh.chunks_of(rank) if e > b]

    def launch(self, out, h, rho0, no_reduction=False):
        import ctypes as C
        import torch
        from . import _capi as capi
        sp = _stream_ptr(out)
        for (_j, b, e) in self.chunks:
            capi.check(capi.lib.dg_density_map_device(self.fh, h, rho0, int(no_reduction), b, e, C.c_void_p(out.data_ptr() + 8 * b), sp))
        return len(self.chunks)

    def step(self, out, h, rho0, no_reduction=False, group=None):
        n = self.launch(out, h, rho0, no_reduction)
        allgather_rows(out, self.sh, group)
        return n


# ---------------------------------------------------------------------------------------------------------------------------
# Slab sharding (bench.py --sharding slab; the node-id chunks above are the default: measured 10.5 vs 10.1 ms per step at 8 GPUs).  Each rank samples whole slow-plane pairs of each of the four node arrays in
# ONE launch (dg_sample_sdf_slab_device): z-slabs of the vertex and x-edge nodes, x-slabs of the y-edge nodes, y-slabs of the z-edge
# nodes.  No partially filled bricks, one launch tail per rank, and the three slab orientations average out the spatially varying
# cost.  A rank therefore owns four contiguous node ranges whose lengths differ slightly between ranks; each node array is
# exchanged with ONE all-gather of equal-sized (padded) slices followed by a copy of the foreign slices to their final offsets.
class SlabSharding:
    def __init__(self, desc, world):
        import ctypes as C
        from . import _capi as capi
        self.world = world
        self.ranges = []                                   # ranges[rank][a] = (l_begin, l_end)
        for r in range(world):
            buf = (C.c_uint64 * 8)()
            capi.check(capi.lib.dg_slab_ranges(C.byref(desc), r, world, buf))
            self.ranges.append([(int(buf[2 * a]), int(buf[2 * a + 1])) for a in range(4)])
        self.maxlen = [max(self.ranges[r][a][1] - self.ranges[r][a][0] for r in range(world)) for a in range(4)]
        n = C.c_uint64()
        capi.check(capi.lib.dg_grid_num_nodes(desc.resolution, C.byref(n)))
        self.n = n.value
        self.padded = self.n + max(self.maxlen)            # so that every [l_begin, l_begin + maxlen) slice stays in bounds

    def covers_exactly_once(self):
        import numpy as np
        seen = np.zeros(self.n, np.int32)
        for r in range(self.world):
            for (b, e) in self.ranges[r]:
                seen[b:e] += 1
        return bool((seen == 1).all())


def allgather_slabs(full, sh, group=None, scratch=None):
    """full: 1-D tensor of sh.padded elements in which this rank has filled its own four ranges; afterwards all ranges are filled."""
    import torch
    import torch.distributed as dist
    if sh.world == 1:
        return
    rank = dist.get_rank(group)
    for a in range(4):
        ml = sh.maxlen[a]
        if ml == 0:
            continue
        b_me = sh.ranges[rank][a][0]
        tmp = scratch[a] if scratch is not None else torch.empty(sh.world * ml, dtype=full.dtype, device=full.device)
        src = full[b_me:b_me + ml]
        dist.all_gather_into_tensor(tmp, src.clone() if full.device.type == "cpu" else src, group=group)
        for r in range(sh.world):
            if r == rank:
                continue
            b, e = sh.ranges[r][a]
            if e > b:
                full[b:e].copy_(tmp[r * ml:r * ml + (e - b)])


class SlabSdfSampler:
    """one rank of the slab-sharded node loop: launch() = this rank's single kernel, step() = launch + exchange"""

    def __init__(self, md, desc, rank, world):
        self.md, self.desc, self.rank, self.world = md, desc, rank, world
        self.sh = SlabSharding(desc, world)
        self.scratch = None

    def launch(self, full, sign=1.0):
        import ctypes as C
        import torch
        from . import _capi as capi
        capi.check(capi.lib.dg_sample_sdf_slab_device(self.md.handle, C.byref(self.desc), sign, self.rank, self.world,
                                                      C.c_void_p(full.data_ptr()), _stream_ptr(full)))
        return 1

    def step(self, full, sign=1.0, group=None):
        import torch
        if self.scratch is None and self.world > 1:
            self.scratch = [torch.empty(self.world * ml, dtype=full.dtype, device=full.device) for ml in self.sh.maxlen]
        self.launch(full, sign)
        allgather_slabs(full, self.sh, group, self.scratch)
        return 1


# ---------------------------------------------------------------------------------------------------------------------------
# Interleaved slab sharding: plane pairs dealt round-robin (pair p -> rank p % world).  One launch per rank, no masked
# bricks, near-perfect balance; ONE in-place all-gather of equal slots; one unpack kernel into the reference's node order.
def allgather_slots(slots, slot_elems, rank, world, group=None):
    """the exchange step: ONE in-place all-gather of the equal-sized slots (rank r's slot = slots[r*slot_elems : (r+1)*slot_elems])"""
    import torch.distributed as dist
    if world > 1:
        dist.all_gather_into_tensor(slots, slots[rank * slot_elems:(rank + 1) * slot_elems], group=group)


def interleaved_node_slots(desc, world, l_begin, l_end):
    """host statement of the layout (dg_interleaved_node_slots): (part, position-in-slot) of nodes [l_begin, l_end) as numpy arrays"""
    import ctypes as C
    import numpy as np
    from . import _capi as capi
    part = np.empty(l_end - l_begin, np.uint32); pos = np.empty(l_end - l_begin, np.uint64)
    capi.check(capi.lib.dg_interleaved_node_slots(C.byref(desc), world, l_begin, l_end, capi.ptr(part, capi.U32P), capi.ptr(pos, capi.U64P)))
    return part, pos


class InterleavedSdfSampler:
    """splits > 1 (experimental, default 1): the deal is made over world * splits parts and rank r takes parts r*splits .. r*splits +
    splits - 1, one launch each on its own stream -- the launches' ~2 ms tails overlap as they do for the node-id chunks, and the
    rank's slots stay contiguous, so the exchange is still ONE all-gather.  world * splits <= 16."""

    def __init__(self, md, desc, rank, world, splits=1):
        import ctypes as C
        from . import _capi as capi
        if splits < 1 or world * splits > 16:
            raise ValueError("world * splits must be in 1..16")
        self.md, self.desc, self.rank, self.world, self.splits = md, desc, rank, world, splits
        self.parts = world * splits
        n = C.c_uint64(); capi.check(capi.lib.dg_grid_num_nodes(desc.resolution, C.byref(n)))
        se = C.c_uint64(); capi.check(capi.lib.dg_interleaved_slot_elems(C.byref(desc), self.parts, C.byref(se)))
        self.n, self.slot = n.value, se.value              # slot = one PART's slot; a rank owns `splits` consecutive ones
        self.slots = None
        self.streams = None

    class _Sh:                                             # the attribute bench.py reads from every sampler
        def __init__(self, padded):
            self.padded = padded

    @property
    def sh(self):
        return InterleavedSdfSampler._Sh(self.n)

    def _buffers(self, full):
        import torch
        if self.slots is None:
            self.slots = torch.empty(self.parts * self.slot, dtype=full.dtype, device=full.device)
            self.streams = _side_streams(full, self.splits if self.splits > 1 else 0)

    def launch(self, full, sign=1.0):
        import ctypes as C
        import torch
        from . import _capi as capi
        self._buffers(full)
        cur = torch.cuda.current_stream() if self.streams else None
        for st in self.streams:
            st.wait_stream(cur)
        for k in range(self.splits):
            part = self.rank * self.splits + k
            sp = C.c_void_p(self.streams[k].cuda_stream) if self.streams else _stream_ptr(full)
            capi.check(capi.lib.dg_sample_sdf_interleaved_device(self.md.handle, C.byref(self.desc), sign, part, self.parts,
                                                                 C.c_void_p(self.slots.data_ptr() + 8 * part * self.slot), sp))
        for st in self.streams:
            cur.wait_stream(st)
        return self.splits

    def step(self, full, sign=1.0, group=None):
        import ctypes as C
        import torch
        from . import _capi as capi
        self.launch(full, sign)
        allgather_slots(self.slots, self.slot * self.splits, self.rank, self.world, group)
        capi.check(capi.lib.dg_interleaved_unpack_device(C.byref(self.desc), self.parts, C.c_void_p(self.slots.data_ptr()),
                                                         C.c_void_p(full.data_ptr()), _stream_ptr(full)))
        return self.splits + 1
