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


MCU_HEIGHT = {"444": 8, "422": 8, "420": 16}


def band_plan(height, subsampling, n_bands):
    """[ext] Splits one frame into `n_bands` horizontal bands of whole MCU rows, so that the bands can
    be encoded independently (on different GPUs) with one restart interval per MCU row and joined
    into ONE baseline JPEG.  Returns a list of (y0, y1, mcu_rows_before) with
    pixel rows [y0, y1); bands that would be empty are dropped."""
    mh = MCU_HEIGHT[str(subsampling)]
    rows = (int(height) + mh - 1) // mh
    n_bands = max(1, min(int(n_bands), rows))
    out = []
    for b in range(n_bands):
        r0, r1 = rows * b // n_bands, rows * (b + 1) // n_bands
        if r1 > r0:
            out.append((r0 * mh, min(r1 * mh, int(height)), r0))
    return out


def _segments(jpg):
    """(marker, offset of the 0xFF, offset past the segment) of every marker segment up to and including SOS."""
    pos = 2
    while True:
        if jpg[pos] != 0xFF:
            raise ValueError("bad JPEG marker structure")
        marker, ln = jpg[pos + 1], (jpg[pos + 2] << 8) | jpg[pos + 3]
        yield marker, pos, pos + 2 + ln
        pos += 2 + ln
        if marker == 0xDA:
            return


def _header_end(jpg):
    """Offset of the first scan byte: walks the marker segments up to and including SOS."""
    end = 0
    for _, _, end in _segments(jpg):
        pass
    return end


def assemble_bands(band_files, plan, height):
    """Joins the JFIF files of the bands of `band_plan` (each encoded with restart_interval = one
    MCU row and restart marker phase = mcu_rows_before % 8) into the file of the whole frame: the
    first band's header with the frame's height, then every band's scan with an RSTm marker in
    place of the band's EOI, m continuing the count of MCU rows modulo 8."""
    if len(band_files) != len(plan):
        raise ValueError("need one file per band")
    first = bytes(band_files[0])
    end = _header_end(first)
    header = bytearray(first[:end])
    # (a search for the bytes FF C0 could land inside a quantisation table given as explicit `tables`)
    sof = next(pos for marker, pos, _ in _segments(first) if marker == 0xC0)
    header[sof + 5] = (int(height) >> 8) & 0xFF                 # SOF0: Lf(2) P(1) Y(2) X(2)
    header[sof + 6] = int(height) & 0xFF
    parts = [bytes(header)]
    for i, (f, (y0, y1, before)) in enumerate(zip(band_files, plan)):
        f = bytes(f)
        if f[-2:] != b"\xFF\xD9":
            raise ValueError("band %d does not end with EOI" % i)
        parts.append(f[_header_end(f):-2])
        if i + 1 < len(plan):
            rows_so_far = plan[i + 1][2]
            parts.append(bytes([0xFF, 0xD0 + ((rows_so_far - 1) & 7)]))
    parts.append(b"\xFF\xD9")
    return b"".join(parts)


class JpegEncoder:
    """RGB frames -> baseline JFIF files on one GPU.

    subsampling: "444" | "422" | "420"; quality: 1..100 ([ext] IJG scaling; 85
    reproduces the model's tables enc_frame.py:141-159) or explicit `tables=(qy, qc)`;
    zrl: "eager" (gateware run-length rule, rlecore.py:101-105) | "standard".
    dct: "model" (float64 model, the parity path) | "fixed" ([ext] the gateware's fixed-point DCT
    arithmetic, new_dct6.py:171-396 corrected; <= 4 LSB per coefficient from the model, not parity).
    restart_interval: [ext] MCUs per restart interval (0 = none): DRI + RSTm markers, DC predictors
    restart; the intervals are independent byte-aligned streams.
    layout: [ext] "rgb" | "bgr" | "rgba" | "bgra": the pixel layout of the frames handed to the encoder
    (last dimension 3 or 4); K1 reads them in place, the files are those of the RGB frames.
    """

    def __init__(self, subsampling="444", quality=85, zrl="eager", tables=None, out_cap=None, dct="model",
                 restart_interval=0, layout="rgb"):
        if str(subsampling) not in _lib.SUB:
            raise ValueError("subsampling must be '444', '422' or '420'")
        if zrl not in _lib.ZRL:
            raise ValueError("zrl must be 'eager' or 'standard'")
        if dct not in _lib.DCT_MODE:
            raise ValueError("dct must be 'model' (float64 model, the parity path) or 'fixed' ([ext] gateware arithmetic)")
        if layout not in ("rgb", "bgr", "rgba", "bgra"):
            raise ValueError("layout must be 'rgb', 'bgr', 'rgba' or 'bgra'")
        self.layout = layout
        self.bpp = 4 if layout in ("rgba", "bgra") else 3
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

    MAX_WORKSPACES = 4

    # ------------------------------------------------------------------ helpers
    def desc_for(self, rgb):
        if rgb.dtype != torch.uint8 or rgb.dim() not in (3, 4) or rgb.shape[-1] != self.bpp:
            raise ValueError("%s frames must be uint8 (H, W, %d) or (N, H, W, %d)" % (self.layout, self.bpp, self.bpp))
        n = 1 if rgb.dim() == 3 else rgb.shape[0]
        if n == 0:
            raise ValueError("empty batch")
        h, w = rgb.shape[-3], rgb.shape[-2]
        return _lib.make_desc(h, w, self.subsampling, n, layout=self.layout)

    def _cap(self, desc):
        return int(self.out_cap or _lib.lib().b2j_default_out_cap(ctypes.byref(desc)))

    def _workspace(self, desc, device, stream=None):
        """Workspace of one in-flight call.  It holds the ticket, the chunk status words and the packed
        stream, so two calls that may overlap must not share it: it is keyed by the stream as well
        as by the shape.  A few are kept; before an older one is dropped the device is synchronised
        (a kernel on another stream may still be using it)."""
        cap = self._cap(desc)
        key = (desc.height, desc.width, desc.n_frames, cap, str(device), self._opts, stream_ptr(stream) or 0)
        ws = self._ws.get(key)
        if ws is None:
            nbytes = _lib.lib().b2j_encode_workspace_bytes_opts(ctypes.byref(desc), cap, self._opts)
            if nbytes == 0:
                raise _lib.B2JError(-1, _lib.last_error())
            if len(self._ws) >= self.MAX_WORKSPACES:
                torch.cuda.synchronize(device)
                self._ws = {}
            ws = torch.empty((nbytes,), dtype=torch.uint8, device=device)
            self._ws[key] = ws
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
        sizes[i] < 0 means slot i was too small (|sizes[i]| bytes are needed).
        Calls on different streams use different workspaces and may overlap; calls on one stream
        are ordered by the stream.  `out` / `sizes`, when given, must have exactly the shapes above."""
        require_cuda()
        if not rgb.is_cuda:
            raise ValueError("rgb must be a CUDA tensor")
        rgb = rgb.contiguous()
        desc = self.desc_for(rgb)
        ws, cap = self._workspace(desc, rgb.device, stream)
        if out is None:
            out = torch.empty((desc.n_frames, cap), dtype=torch.uint8, device=rgb.device)
        elif (out.dtype != torch.uint8 or tuple(out.shape) != (desc.n_frames, cap) or out.device != rgb.device
              or not out.is_contiguous()):
            raise ValueError("out must be a contiguous uint8 tensor of shape (%d, %d) on %s" % (desc.n_frames, cap, rgb.device))
        if sizes is None:
            sizes = torch.empty((desc.n_frames,), dtype=torch.int64, device=rgb.device)
        elif (sizes.dtype != torch.int64 or tuple(sizes.shape) != (desc.n_frames,) or sizes.device != rgb.device
              or not sizes.is_contiguous()):
            raise ValueError("sizes must be a contiguous int64 tensor of shape (%d,) on %s" % (desc.n_frames, rgb.device))
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
        # only the bytes every file really has cross PCIe (a slot is several times a typical file)
        n = [int(v) for v in sizes]
        flat = torch.cat([out[i, :n[i]] for i in range(out.shape[0])]).cpu().numpy().tobytes()
        files, pos = [], 0
        for k in n:
            files.append(flat[pos:pos + k])
            pos += k
        return files

    def band_encoder(self, width, mcu_rows_before):
        """The encoder for one band of `band_plan`: same tables and policies, one restart interval
        per MCU row, restart markers counted from the band's position in the frame."""
        if self.restart_interval:
            raise ValueError("band splitting chooses the restart interval itself")
        mcus_x = _lib.geometry(_lib.make_desc(8, width, self.subsampling, 1)).mcus_x
        enc = JpegEncoder(self.subsampling, quality=self.quality or 85, zrl=self.zrl, dct=self.dct, restart_interval=mcus_x,
                          tables=None if self.quality else (self.qy, self.qc), out_cap=self.out_cap)
        enc._opts |= (int(mcu_rows_before) & 7) << 12            # B2J_OPT_RST_PHASE
        return enc

    def encode_frame_bands(self, rgb, n_bands):
        """[ext] One frame (H, W, 3) encoded as `n_bands` independent bands on this GPU and joined:
        byte-identical to encoding the whole frame with restart_interval = MCUs per row.  The
        multi-GPU form is ShardedJpegEncoder.encode_frame_split."""
        if rgb.dim() != 3:
            raise ValueError("encode_frame_bands takes one frame (H, W, 3)")
        h, w = rgb.shape[0], rgb.shape[1]
        plan = band_plan(h, self.subsampling, n_bands)
        files = []
        for (y0, y1, before) in plan:
            files.append(self.band_encoder(w, before).encode(rgb[y0:y1][None])[0])
        return assemble_bands(files, plan, h)

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
        fpc = int(frames_per_chunk or max(1, min(n, (64 << 20) // max(1, desc.height * desc.width * self.bpp))))
        cap = self._cap(_lib.make_desc(desc.height, desc.width, self.subsampling, 1))
        key = (desc.height, desc.width, fpc, cap, torch.cuda.current_device())
        enc = self._host.get(key)
        if enc is None:
            self.close()
            handle = ctypes.c_void_p()
            d1 = _lib.make_desc(desc.height, desc.width, self.subsampling, fpc, layout=self.layout)
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


# ---------------------------------------------------------------------- multi-GPU (one process, one call)
class MultiGpuJpegEncoder:
    """[ext] One host batch, several GPUs of one box, ONE call (C entry b2j_multi_encoder_*): frames
    shard_range(N, G, g) are encoded on devices[g], each device driven by its own host thread inside
    the library; no exchange between devices.  The unit that shards is the per-image loop
    test/model/enc_frame.py:203-238."""

    def __init__(self, devices, height, width, subsampling="420", quality=75, zrl="eager", dct="model", restart_interval=0,
                 out_cap=None, frames_per_chunk=16):
        require_cuda()
        self.devices = [int(d) for d in devices]
        if not self.devices:
            raise ValueError("at least one device")
        self.subsampling = str(subsampling)
        self.height, self.width = int(height), int(width)
        opts = _lib.ZRL[zrl] | _lib.DCT_MODE[dct] | _lib.opt_restart(restart_interval)
        d1 = _lib.make_desc(self.height, self.width, self.subsampling, int(frames_per_chunk))
        self.out_cap = int(out_cap or _lib.lib().b2j_default_out_cap(ctypes.byref(_lib.make_desc(self.height, self.width, self.subsampling, 1))))
        arr = (ctypes.c_int32 * len(self.devices))(*self.devices)
        self._h = ctypes.c_void_p()
        _lib.check(_lib.lib().b2j_multi_encoder_create(ctypes.byref(self._h), arr, len(self.devices), ctypes.byref(d1),
                                                        int(quality), opts, self.out_cap, int(frames_per_chunk)))

    def _check(self, rgb_host):
        if rgb_host.is_cuda or rgb_host.dtype != torch.uint8 or rgb_host.dim() != 4 or \
                tuple(rgb_host.shape[1:]) != (self.height, self.width, 3) or not rgb_host.is_contiguous():
            raise ValueError("rgb_host must be a contiguous CPU uint8 tensor (N, %d, %d, 3)" % (self.height, self.width))

    def encode_host(self, rgb_host, out_host=None, sizes_host=None):
        """(N, H, W, 3) CPU uint8 (pinned for speed) -> (out_host (N, cap) uint8, sizes_host (N,) int64)."""
        self._check(rgb_host)
        n = rgb_host.shape[0]
        if out_host is None:
            out_host = torch.empty((n, self.out_cap), dtype=torch.uint8).pin_memory()
        if sizes_host is None:
            sizes_host = torch.empty((n,), dtype=torch.int64)
        if tuple(out_host.shape) != (n, self.out_cap) or out_host.dtype != torch.uint8 or not out_host.is_contiguous() or \
                tuple(sizes_host.shape) != (n,) or sizes_host.dtype != torch.int64:
            raise ValueError("out_host must be (N, %d) uint8 and sizes_host (N,) int64" % self.out_cap)
        _lib.check(_lib.lib().b2j_multi_encoder_encode_host(self._h, rgb_host.data_ptr(), n, out_host.data_ptr(),
                                                             sizes_host.data_ptr()))
        return out_host, sizes_host

    def h2d_seconds(self, rgb_host, repeats=3):
        """Wall-clock seconds of ONE pass of only the host-to-device copies encode_host would issue
        (all devices concurrently): the copy ceiling of the box for this batch."""
        self._check(rgb_host)
        sec = ctypes.c_double()
        _lib.check(_lib.lib().b2j_multi_encoder_h2d_probe(self._h, rgb_host.data_ptr(), rgb_host.shape[0], int(repeats),
                                                           ctypes.byref(sec)))
        return float(sec.value)

    def close(self):
        if self._h:
            _lib.lib().b2j_multi_encoder_destroy(self._h)
            self._h = ctypes.c_void_p()

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

    def encode_frame_split(self, rgb_frame, dst=0, group=None):
        """[ext] ONE frame across all ranks (a single 8K frame cannot use 8 GPUs without
        restart markers): rank r encodes band r of band_plan(H, subsampling, world_size) on its
        GPU; the finished band files are gathered on `dst` (output only, no collective on the data
        path) and joined into one baseline JPEG.  Returns the file on `dst`, None elsewhere."""
        import torch.distributed as dist
        h, w = rgb_frame.shape[0], rgb_frame.shape[1]
        plan = band_plan(h, self.encoder.subsampling, self.world_size)
        mine = None
        if self.rank < len(plan):
            y0, y1, before = plan[self.rank]
            mine = self.encoder.band_encoder(w, before).encode(rgb_frame[y0:y1][None])[0]
        gathered = [None] * self.world_size if dist.get_rank(group) == dst else None
        dist.gather_object(mine, gathered, dst=dst, group=group)
        if gathered is None:
            return None
        return assemble_bands([g for g in gathered if g is not None], plan, h)

    def encode_local(self, rgb_all_or_local, n_frames=None, already_local=False):
        """Encodes this rank's shard on its GPU; returns (out, sizes) device tensors of the shard."""
        if already_local:
            return self.encoder.encode_device(rgb_all_or_local)
        lo, hi = self.local_range(n_frames if n_frames is not None else rgb_all_or_local.shape[0])
        return self.encoder.encode_device(rgb_all_or_local[lo:hi].cuda(non_blocking=True))
