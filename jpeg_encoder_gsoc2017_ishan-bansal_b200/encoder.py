"""[ext] Whole-frame encoder: the composition the reference writes out in
test/model/enc_frame.py:203-238 (rgb2ycbcr -> dct -> quantize -> zigzag -> DC
difference -> RLE -> Huffman), run as the fused CUDA path K1..K4 behind the C
ABI of include/b2j.h, plus frame sharding across the GPUs of one box.

The encode path shards trivially (frames are independent; the only sequential
dependencies - DC prediction and bit offsets - live inside one frame), so there
is no collective on the hot path: each rank encodes a contiguous range of
frames and only byte counts / finished streams are gathered for output.
"""
import ctypes

import torch

from . import _lib
from .common import require_cuda, stream_ptr


def shard_range(n_frames, world_size, rank):
    """Contiguous frame range [lo, hi) of `rank`: [g*B/G, (g+1)*B/G)."""
    if not 0 <= rank < world_size:
        raise ValueError("rank out of range")
    return (n_frames * rank) // world_size, (n_frames * (rank + 1)) // world_size


class JpegEncoder:
    """RGB frames -> baseline JFIF files on one GPU.

    subsampling: "444" | "422" | "420"; quality: 1..100 ([ext] IJG scaling; 85
    reproduces the model's tables enc_frame.py:141-159) or explicit `tables=(qy, qc)`;
    zrl: "eager" (gateware run-length rule, rlecore.py:101-105) | "standard".
    dct: "model" (float64 model, the parity path) | "fixed" ([ext] the gateware's fixed-point DCT
    arithmetic, new_dct6.py:171-396 corrected; <= 4 LSB per coefficient from the model, not parity).
    restart_interval: [ext] MCUs per restart interval (0 = none): DRI + RSTm markers, DC predictors
    restart; the intervals are independent byte-aligned streams.
    """

    def __init__(self, subsampling="444", quality=85, zrl="eager", tables=None, out_cap=None, dct="model",
                 restart_interval=0):
        if str(subsampling) not in _lib.SUB:
            raise ValueError("subsampling must be '444', '422' or '420'")
        if zrl not in _lib.ZRL:
            raise ValueError("zrl must be 'eager' or 'standard'")
        if dct not in _lib.DCT_MODE:
            raise ValueError("dct must be 'model' (float64 model, the parity path) or 'fixed' ([ext] gateware arithmetic)")
        self.subsampling = str(subsampling)
        self.zrl = zrl
        self.dct = dct
        self.restart_interval = int(restart_interval)
        self._opts = _lib.ZRL[zrl] | _lib.DCT_MODE[dct] | _lib.opt_restart(restart_interval)
        if tables is not None:
            self.quality = 0
            self.qy, self.qc = [int(v) for v in tables[0]], [int(v) for v in tables[1]]
            if len(self.qy) != 64 or len(self.qc) != 64 or min(self.qy + self.qc) < 1 or max(self.qy + self.qc) > 255:
                raise ValueError("tables must be two lists of 64 divisors in 1..255")
        else:
            if not 1 <= int(quality) <= 100:
                raise ValueError("quality must be in 1..100")
            self.quality = int(quality)
            self.qy, self.qc = _lib.quant_tables(self.quality)
        self.out_cap = out_cap
        self._ws = {}
        self._host = {}

    # ------------------------------------------------------------------ helpers
    def desc_for(self, rgb):
        if rgb.dtype != torch.uint8 or rgb.dim() not in (3, 4) or rgb.shape[-1] != 3:
            raise ValueError("rgb must be uint8 (H, W, 3) or (N, H, W, 3)")
        n = 1 if rgb.dim() == 3 else rgb.shape[0]
        if n == 0:
            raise ValueError("empty batch")
        h, w = rgb.shape[-3], rgb.shape[-2]
        return _lib.make_desc(h, w, self.subsampling, n)

    def _cap(self, desc):
        return int(self.out_cap or _lib.lib().b2j_default_out_cap(ctypes.byref(desc)))

    def _workspace(self, desc, device):
        cap = self._cap(desc)
        key = (desc.height, desc.width, desc.n_frames, cap, str(device), self._opts)
        ws = self._ws.get(key)
        if ws is None:
            nbytes = _lib.lib().b2j_encode_workspace_bytes_opts(ctypes.byref(desc), cap, self._opts)
            if nbytes == 0:
                raise _lib.B2JError(-1, _lib.last_error())
            ws = torch.empty((nbytes,), dtype=torch.uint8, device=device)
            self._ws = {key: ws}            # keep one shape resident
        return ws, cap

    # ------------------------------------------------------------------ device path
    def coefficients(self, rgb, variant="auto", stream=None):
        """K1 only: quantised zig-zag coefficients (N, blocks_per_frame, 64) int16, scan order."""
        require_cuda()
        if not rgb.is_cuda:
            raise ValueError("rgb must be a CUDA tensor")
        rgb = rgb.contiguous()
        desc = self.desc_for(rgb)
        g = _lib.geometry(desc)
        coefs = torch.empty((desc.n_frames, g.blocks_per_frame, 64), dtype=torch.int16, device=rgb.device)
        with torch.cuda.device(rgb.device):
            _lib.check(_lib.lib().b2j_fused_block(rgb.data_ptr(), ctypes.byref(desc), _lib.table_arg(self.qy),
                                                   _lib.table_arg(self.qc), coefs.data_ptr(),
                                                   _lib.FUSED[variant] | _lib.DCT_MODE[self.dct], stream_ptr(stream)))
        return coefs

    def encode_device(self, rgb, out=None, sizes=None, stream=None):
        """K1..K4 on device buffers, asynchronous.  Returns (out (N, cap) uint8, sizes (N,) int64);
        sizes[i] < 0 means slot i was too small (|sizes[i]| bytes are needed)."""
        require_cuda()
        if not rgb.is_cuda:
            raise ValueError("rgb must be a CUDA tensor")
        rgb = rgb.contiguous()
        desc = self.desc_for(rgb)
        ws, cap = self._workspace(desc, rgb.device)
        if out is None:
            out = torch.empty((desc.n_frames, cap), dtype=torch.uint8, device=rgb.device)
        if sizes is None:
            sizes = torch.empty((desc.n_frames,), dtype=torch.int64, device=rgb.device)
        with torch.cuda.device(rgb.device):
            _lib.check(_lib.lib().b2j_encode(rgb.data_ptr(), ctypes.byref(desc), self.quality, _lib.table_arg(self.qy),
                                              _lib.table_arg(self.qc), self._opts, ws.data_ptr(),
                                              ws.numel(),
                                              out.data_ptr(), cap, sizes.data_ptr(), stream_ptr(stream)))
        return out, sizes

    def encode(self, rgb):
        """Convenience: CUDA or CPU uint8 frames -> list of `bytes` (one JFIF file per frame)."""
        require_cuda()
        dev_rgb = rgb if rgb.is_cuda else rgb.cuda()
        out, sizes = self.encode_device(dev_rgb)
        sizes = sizes.cpu()
        if int(sizes.min()) < 0:
            raise _lib.B2JError(-4, "a frame needs %d bytes but the slot holds %d" % (-int(sizes.min()), out.shape[1]))
        host = out.cpu()
        return [bytes(host[i, :int(sizes[i])].numpy().tobytes()) for i in range(out.shape[0])]

    def launch_count(self, desc):
        return int(_lib.lib().b2j_encode_launch_count(ctypes.byref(desc)))

    # ------------------------------------------------------------------ host-buffer path (end to end)
    def encode_host(self, rgb_host, out_host=None, sizes_host=None, frames_per_chunk=None):
        """CPU uint8 frames (pinned or pageable) -> (out_host (N, cap) uint8, sizes_host (N,) int64).
        H2D copies, kernels and D2H copies of consecutive chunks overlap inside the library."""
        require_cuda()
        if rgb_host.is_cuda:
            raise ValueError("encode_host takes a CPU tensor")
        rgb_host = rgb_host.contiguous()
        desc = self.desc_for(rgb_host)
        n = desc.n_frames
        fpc = int(frames_per_chunk or max(1, min(n, (64 << 20) // max(1, desc.height * desc.width * 3))))
        cap = self._cap(_lib.make_desc(desc.height, desc.width, self.subsampling, 1))
        key = (desc.height, desc.width, fpc, cap, torch.cuda.current_device())
        enc = self._host.get(key)
        if enc is None:
            self.close()
            handle = ctypes.c_void_p()
            d1 = _lib.make_desc(desc.height, desc.width, self.subsampling, fpc)
            if self.quality == 0:
                raise ValueError("encode_host needs a quality (explicit tables are not supported on this path)")
            _lib.check(_lib.lib().b2j_encoder_create(ctypes.byref(handle), ctypes.byref(d1), self.quality, self._opts,
                                                      cap, fpc))
            enc = handle
            self._host = {key: enc}
        if out_host is None:
            out_host = torch.empty((n, cap), dtype=torch.uint8).pin_memory()
        if sizes_host is None:
            sizes_host = torch.empty((n,), dtype=torch.int64)
        _lib.check(_lib.lib().b2j_encoder_encode_host(enc, rgb_host.data_ptr(), n, out_host.data_ptr(), sizes_host.data_ptr()))
        return out_host, sizes_host

    def close(self):
        for enc in self._host.values():
            _lib.lib().b2j_encoder_destroy(enc)
        self._host = {}

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---------------------------------------------------------------------- multi-GPU (one process per GPU)
def gather_sizes(sizes_local, n_frames, group=None):
    """All ranks learn every frame's byte count.  The one (optional) collective of
    the path: an all_gather of int64 counts, off the hot path."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    per = max(shard_range(n_frames, world, r)[1] - shard_range(n_frames, world, r)[0] for r in range(world))
    pad = torch.full((per,), -1, dtype=torch.int64, device=sizes_local.device)
    pad[:sizes_local.numel()] = sizes_local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    out = []
    for r in range(world):
        lo, hi = shard_range(n_frames, world, r)
        out.append(bufs[r][:hi - lo])
    del rank
    return torch.cat(out)


def gather_streams(streams_local, n_frames, dst=0, group=None):
    """Collects the finished JFIF byte strings of all shards on rank `dst` (output only)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    gathered = [None] * world if dist.get_rank(group) == dst else None
    dist.gather_object(list(streams_local), gathered, dst=dst, group=group)
    if gathered is None:
        return None
    out = [s for part in gathered for s in part]
    if len(out) != n_frames:
        raise RuntimeError("gathered %d streams, expected %d" % (len(out), n_frames))
    return out


class ShardedJpegEncoder:
    """One process per GPU; rank r encodes frames shard_range(N, world, r)."""

    def __init__(self, world_size, rank, **encoder_kwargs):
        self.world_size, self.rank = int(world_size), int(rank)
        self.encoder = JpegEncoder(**encoder_kwargs)

    def local_range(self, n_frames):
        return shard_range(n_frames, self.world_size, self.rank)

    def encode_local(self, rgb_all_or_local, n_frames=None, already_local=False):
        """Encodes this rank's shard on its GPU; returns (out, sizes) device tensors of the shard."""
        if already_local:
            return self.encoder.encode_device(rgb_all_or_local)
        lo, hi = self.local_range(n_frames if n_frames is not None else rgb_all_or_local.shape[0])
        return self.encoder.encode_device(rgb_all_or_local[lo:hi].cuda(non_blocking=True))
