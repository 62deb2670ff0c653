"""Channel sharding across ranks (one process per GPU).

Channels are independent streams, so the decode itself partitions with no data-path collective: rank r of W owns a
contiguous channel range (`channel_range`). The only exchange step the path can have is transport when the I/O is
single-homed (SURVEY.md §8 e): one rank reads every channel's file and owns the sinks. `SingleHomedTransport` is that
step — IQ super-blocks scattered from the ingest rank in the FILE's sample format (2-8 bytes per sample, decoded on
the owning GPU), sink-format audio gathered back — over torch.distributed (NCCL over NVLink/NVSwitch on the GPU box,
gloo in the CPU tests). `ShardedDecoder` puts a decoder of this package between the two.
No torch import at module load: the decoder package itself stays torch-free.
"""


def channel_range(rank, world, total_channels):
    """Contiguous, balanced partition: the first (total % world) ranks get one extra channel."""
    if not (0 <= rank < world) or total_channels < 0:
        raise ValueError("bad rank/world/total")
    base, extra = divmod(total_channels, world)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo, hi


class SingleHomedTransport:
    """Row-wise scatter / gather between the ingest rank and the ranks that own the channels.

    Rows are channels. Every rank receives `rows_per_rank = ceil(total / world)` rows (the trailing rows of the
    last shards are padding when the partition is uneven), so both collectives are the plain equal-size
    scatter / gather that NCCL implements as one grouped send/recv."""

    def __init__(self, total_channels, group=None, root=0):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.root = root
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.total = int(total_channels)
        self.lo, self.hi = channel_range(self.rank, self.world, self.total)
        self.rows_per_rank = -(-self.total // self.world) if self.total else 0

    @property
    def n_local(self):
        return self.hi - self.lo

    def _root_global(self):
        return self.dist.get_global_rank(self.group, self.root) if self.group is not None else self.root

    def scatter_rows(self, full, row_elems, dtype, device):
        """`full`: on the root a [total, row_elems] tensor of `dtype` on `device`, elsewhere None.
        Returns this rank's [rows_per_rank, row_elems] shard (rows >= n_local are padding)."""
        import torch
        recv = torch.empty((self.rows_per_rank, row_elems), dtype=dtype, device=device)
        parts = None
        if self.rank == self.root:
            assert full.shape == (self.total, row_elems) and full.dtype == dtype
            parts = []
            for r in range(self.world):
                lo, hi = channel_range(r, self.world, self.total)
                if hi - lo == self.rows_per_rank:
                    parts.append(full[lo:hi].contiguous())
                else:
                    pad = torch.zeros((self.rows_per_rank, row_elems), dtype=dtype, device=device)
                    pad[:hi - lo] = full[lo:hi]
                    parts.append(pad)
        self.dist.scatter(recv, parts, src=self._root_global(), group=self.group)
        return recv

    def gather_rows(self, local):
        """`local`: [rows_per_rank, n] on every rank. Returns [total, n] on the root, None elsewhere.
        Rows travel as bytes: int16 (the 16-bit sink format) is not a dtype NCCL or gloo reduce / move natively."""
        import torch
        assert local.dim() == 2 and local.shape[0] == self.rows_per_rank
        dtype, n = local.dtype, local.shape[1]
        if n == 0:
            return torch.empty((self.total, 0), dtype=dtype, device=local.device) if self.rank == self.root else None
        raw = local.contiguous().view(torch.uint8)
        parts = None
        if self.rank == self.root:
            parts = [torch.empty_like(raw) for _ in range(self.world)]
        self.dist.gather(raw, parts, dst=self._root_global(), group=self.group)
        if self.rank != self.root:
            return None
        out = torch.empty((self.total, n), dtype=dtype, device=local.device)
        for r in range(self.world):
            lo, hi = channel_range(r, self.world, self.total)
            out[lo:hi] = parts[r].view(dtype)[:hi - lo]
        return out


class ShardedDecoder:
    """`total_channels` streams decoded by all ranks of the group, I/O on the root rank only.

    make_decoder(n_channels) builds this rank's decoder (FmDecoder / AmDecoder / NbfmDecoder of this package, on
    this rank's GPU) for rows_per_rank channels; padding rows decode silence and are dropped by the gather."""

    def __init__(self, total_channels, make_decoder, group=None, root=0):
        self.tr = SingleHomedTransport(total_channels, group=group, root=root)
        self.dec = make_decoder(max(self.tr.rows_per_rank, 1))

    def process_blocks_from_root(self, raw_root, iq_format, block_len, out_format, squelch_level=0.0, gain=0.5,
                                 device=None):
        """raw_root: on the root a uint8 [total, T * bytes_per_complex_sample] device tensor (the files' own bytes),
        elsewhere None. Returns (audio [total, n] in the sink dtype on the root / None elsewhere, audio_len[n_blocks])."""
        import torch

        from . import _capi
        esz = _capi.IQ_BYTES[iq_format]
        total_in = int(sum(int(b) for b in block_len))
        dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        shard = self.tr.scatter_rows(raw_root, total_in * esz, torch.uint8, dev)
        out_total, _ = self.dec.query_output(block_len)
        tdt = {_capi.OUT_F64: torch.float64, _capi.OUT_F32: torch.float32, _capi.OUT_S16: torch.int16}[out_format]
        local = torch.zeros((self.tr.rows_per_rank, max(out_total, 1)), dtype=tdt, device=dev)
        st = torch.cuda.current_stream(dev)
        lens = self.dec.process_device_io(shard.data_ptr(), iq_format, total_in, block_len, local.data_ptr(),
                                          local.shape[1], out_format=out_format, squelch_level=squelch_level, gain=gain,
                                          stream=st.cuda_stream)
        full = self.tr.gather_rows(local[:, :out_total] if out_total else local[:, :0])
        return full, lens
