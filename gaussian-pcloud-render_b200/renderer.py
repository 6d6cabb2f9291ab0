"""FrameRenderer -- the synchronisation-free forward path over preallocated workspaces (gs_forward_nosync).

The drop-in `GaussianRasterizer` keeps the reference's behaviour of sizing the instance buffers from a host
read-back of num_rendered (rasterizer_impl.cu:281).  A renderer that draws many frames of one cloud -- what
PCML_Render.render does per view (simple_raw_render.py:411-522) -- does not need that: this class keeps the three
workspaces resident, sizes the instance buffers for a capacity with headroom, and only looks at the frame status
(num_rendered / overflow) when the caller synchronises anyway.  On overflow the frame is re-issued once with a
larger capacity.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from diff_gaussian_rasterization import _C


class FrameRenderer:
    SLOTS = 1024  # pinned status slots: one per in-flight frame

    def __init__(self, cloud: dict, width: int, height: int, bg, device, capacity: int = 0, headroom: float = 1.3,
                 tile_rows: Optional[Tuple[int, int]] = None, share: Optional["FrameRenderer"] = None,
                 downsample: int = 1, team_after: int = 0, blend_split: int = 0):
        """blend_split: > 0 = latency mode of the blend (GsScene.blend_split): list walks longer than that many batches
        are finished by a CTA as parallel segments merged associatively -- pixels within ~1e-6 of the exact frame, NOT
        bit-identical; 0 = off (every frame bit-identical to the reference kernels).
        downsample=2: width x height is the (super-sampled) raster size, every output image is the 2x2 box mean
        (3, height/2, width/2) -- the reference caller's bilinear x0.5, done in the blend epilogue.
        team_after: blend scheduling hint (GsScene.team_after; results never change): > 0 = list walks longer than
        that many batches are finished by CTA teams (a latency experiment, see csrc/blend_forward.cu), 0 = library
        default (off), -1 = off."""
        self.L = _C.lib()
        self.team_after = int(team_after)
        self.blend_split = int(blend_split)
        if downsample not in (1, 2) or (downsample == 2 and (int(width) % 2 or int(height) % 2)):
            raise ValueError("downsample must be 1 or 2 (2 needs an even raster size)")
        self.downsample = int(downsample)
        self.dev = torch.device(device)
        self.W, self.H = int(width), int(height)
        if share is not None:  # same cloud already resident on the device: share the attribute tensors
            self.means3D, self.opacities, self.scales = share.means3D, share.opacities, share.scales
            self.rotations, self.shs, self.sh_degree = share.rotations, share.shs, share.sh_degree
        else:
            f32 = lambda t: t.to(self.dev, torch.float32).contiguous()
            self.means3D, self.opacities = f32(cloud["means3D"]), f32(cloud["opacities"])
            self.scales, self.rotations = f32(cloud["scales"]), f32(cloud["rotations"])
            self.sh_degree = int(cloud["sh_degree"])
            # resident layout: only the (degree + 1)^2 coefficients the rasterizer reads are kept on the device -- the
            # reference's pcrender clouds carry 13 coefficients per point and render with degree 1, and at DRAM
            # granularity the per-Gaussian stage would otherwise pull in the 9 unread ones of every row too
            # (160 MB instead of 83 MB per frame at 800 K points).  Same frame bit for bit.
            sh = cloud["shs"]
            need = (self.sh_degree + 1) ** 2
            self.shs = f32(sh[:, :need] if sh.shape[1] > need else sh)
        self.P = int(self.means3D.shape[0])
        self.bg = torch.as_tensor(bg, dtype=torch.float32).to(self.dev)
        self.headroom = headroom
        self.tile_rows = tile_rows
        with torch.cuda.device(self.dev):
            self.geom = torch.empty(self.L.gs_geometry_bytes(self.P), dtype=torch.uint8, device=self.dev)
            self.img = torch.empty(self.L.gs_image_bytes(self.W, self.H), dtype=torch.uint8, device=self.dev)
            self.radii = torch.zeros(self.P, dtype=torch.int32, device=self.dev)
            self.color = torch.zeros((3, self.H // self.downsample, self.W // self.downsample), dtype=torch.float32,
                                     device=self.dev)
            self.status_host = torch.zeros((self.SLOTS, 2), dtype=torch.int64).pin_memory()
        self.capacity = 0
        self.binning = None
        self._reserve(max(int(capacity), 1 << 16))

    def _reserve(self, cap: int) -> None:
        self.capacity = int(cap)
        with torch.cuda.device(self.dev):
            self.binning = torch.empty(self.L.gs_binning_bytes(self.capacity, self.P, self.W, self.H),
                                       dtype=torch.uint8, device=self.dev)

    def _scene(self, view_dev, tile_rows, peer_out=None, extra_passes=None, shard_cull=False):
        viewmatrix, projmatrix, campos, tanx, tany = view_dev
        return _C.make_scene(P=self.P, sh_degree=self.sh_degree, sh_stride=int(self.shs.shape[1]), width=self.W,
                             height=self.H, tan_fovx=float(tanx), tan_fovy=float(tany), scale_modifier=1.0,
                             prefiltered=False, debug=False, background=self.bg, means3D=self.means3D, shs=self.shs,
                             colors_precomp=None, opacities=self.opacities, scales=self.scales,
                             rotations=self.rotations, cov3D_precomp=None, viewmatrix=viewmatrix,
                             projmatrix=projmatrix, campos=campos, tile_rows=tile_rows, peer_out=peer_out,
                             extra_passes=extra_passes, downsample=self.downsample, team_after=self.team_after,
                             shard_cull=shard_cull, blend_split=self.blend_split)

    def upload_view(self, view):
        """host View (scenes.make_view) -> device tensors; done once per camera, outside the frame loop."""
        t = lambda a: torch.from_numpy(a).to(self.dev).contiguous()
        return (t(view.viewmatrix), t(view.projmatrix), t(view.campos), view.tanfovx, view.tanfovy)

    def enqueue(self, view_dev, out_color: Optional[torch.Tensor] = None, tile_rows=None, slot: int = 0,
                peer_out=None, extra_passes=None, shard_cull: bool = False) -> torch.Tensor:
        """Queues one frame on the current stream; no host synchronisation.  Returns the (3,H,W) colour tensor.
        peer_out: device pointers of the (3,H,W) images of all ranks (sharding.PeerFrame): the blend epilogue then
        writes this rank's tile rows into every one of them instead of into out_color.
        extra_passes: up to three (colors (P,3), out (3,H,W)) pairs of contiguous fp32 CUDA tensors blended in the same
        list walk as the frame; each `out` equals a separate forward with colors_precomp = colors.
        shard_cull (with tile_rows): Gaussians that cannot reach the shard's rows are dropped before the per-Gaussian
        stage, which then runs -- like the depth sort and the list passes -- on the survivors only (pixels unchanged;
        `self.radii` is only updated for the survivors)."""
        out = self.color if out_color is None else out_color
        scene = self._scene(view_dev, tile_rows if tile_rows is not None else self.tile_rows, peer_out, extra_passes,
                            shard_cull)
        with torch.cuda.device(self.dev):
            st = torch.cuda.current_stream(self.dev).cuda_stream
            _C._check(self.L.gs_forward_nosync(C.byref(scene), self.geom.data_ptr(), self.binning.data_ptr(),
                                               self.capacity, self.img.data_ptr(), out.data_ptr(),
                                               self.radii.data_ptr(), st), "gs_forward_nosync")
            _C._check(self.L.gs_read_status(self.geom.data_ptr(), self.status_host[slot % self.SLOTS].data_ptr(), st),
                      "gs_read_status")
        return out

    def capture_graph(self, tanfov: Tuple[float, float]) -> None:
        """Captures one frame (memsets, the eleven kernels, the status read-back) into a CUDA graph that reads its
        camera from a fixed 48-float slot of this renderer, so that a frame costs the host one small copy and one
        graph launch instead of ~15 launches -- for small frames (C1: ~0.1 ms of GPU per frame with several frames in
        flight) the Python/launch path is otherwise the bottleneck.  Replay with `enqueue_graph`."""
        with torch.cuda.device(self.dev):
            self._gview = torch.zeros(_C.GS_VIEW_STRIDE, dtype=torch.float32, device=self.dev)
            self._gview[0] = self._gview[5] = self._gview[10] = self._gview[15] = 1.0
            view = (self._gview[0:16], self._gview[16:32], self._gview[32:35], float(tanfov[0]), float(tanfov[1]))
            self.enqueue(view)  # the library's one-time set-up (function attributes, occupancy queries) happens here
            torch.cuda.synchronize(self.dev)
            self._graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph):
                self.enqueue(view, slot=0)

    def enqueue_graph(self, view_row: torch.Tensor) -> torch.Tensor:
        """One frame through the captured graph; `view_row` is a row of ViewBatch.buf (viewmatrix, projmatrix, campos).
        Returns this renderer's colour buffer (overwritten by its next frame); status() reports slot 0."""
        self._gview.copy_(view_row, non_blocking=True)
        self._graph.replay()
        return self.color

    def enqueue_pass(self, view_dev, out_color: torch.Tensor, *, colors_precomp: Optional[torch.Tensor] = None,
                     shs: Optional[torch.Tensor] = None, sh_degree: Optional[int] = None, tile_rows=None) -> torch.Tensor:
        """Another colour pass over the frame this renderer enqueued last (same cloud geometry, same view): only
        the per-Gaussian colours are recomputed and the blend re-run -- preprocess, depth sort and tile lists are
        reused (the reference's caller renders position / RGB / hit-map / normal passes per view).
        Exactly one of colors_precomp (P,3) / shs (P,M,3) is given; the result goes to `out_color` (3,H,W)."""
        if (colors_precomp is None) == (shs is None):
            raise ValueError("give exactly one of colors_precomp / shs")
        viewmatrix, projmatrix, campos, tanx, tany = view_dev
        f32 = lambda t: None if t is None else t.to(self.dev, torch.float32).contiguous()
        cp, sh = f32(colors_precomp), f32(shs)
        scene = _C.make_scene(P=self.P, sh_degree=int(self.sh_degree if sh_degree is None else sh_degree),
                              sh_stride=0 if sh is None else int(sh.shape[1]), width=self.W, height=self.H,
                              tan_fovx=float(tanx), tan_fovy=float(tany), scale_modifier=1.0, prefiltered=False,
                              debug=False, background=self.bg, means3D=self.means3D, shs=sh, colors_precomp=cp,
                              opacities=self.opacities, scales=self.scales, rotations=self.rotations,
                              cov3D_precomp=None, viewmatrix=viewmatrix, projmatrix=projmatrix, campos=campos,
                              tile_rows=tile_rows if tile_rows is not None else self.tile_rows,
                              downsample=self.downsample, team_after=self.team_after)
        with torch.cuda.device(self.dev):
            st = torch.cuda.current_stream(self.dev).cuda_stream
            _C._check(self.L.gs_forward_recolor(C.byref(scene), self.geom.data_ptr(), self.binning.data_ptr(),
                                                self.img.data_ptr(), out_color.data_ptr(), st), "gs_forward_recolor")
        self._keep = (cp, sh)  # keep the converted tensors alive until the stream has consumed them
        return out_color

    def status(self, slot: int = 0) -> Tuple[int, int, int]:
        """(num_rendered, num_visible, code) of the frame enqueued with `slot`; call after synchronising the stream."""
        nr = int(self.status_host[slot % self.SLOTS, 0])
        w = int(self.status_host[slot % self.SLOTS, 1])
        return nr, w & 0xffffffff, C.c_int32((w >> 32) & 0xffffffff).value

    def render(self, view_dev, out_color: Optional[torch.Tensor] = None, tile_rows=None) -> torch.Tensor:
        """Frame + synchronise + overflow handling (grows the instance buffers and re-renders if needed)."""
        out = self.enqueue(view_dev, out_color, tile_rows)
        torch.cuda.current_stream(self.dev).synchronize()
        nr, _, code = self.status()
        if code == -4:
            self._reserve(int(nr * self.headroom) + 1024)
            out = self.enqueue(view_dev, out_color, tile_rows)
            torch.cuda.current_stream(self.dev).synchronize()
            nr, _, code = self.status()
        if code != 0:
            raise RuntimeError(f"render failed: {_C.GS_ERRORS.get(code, code)}")
        return out

    def calibrate(self, views_dev) -> int:
        """Renders each view once and sizes the instance buffers for the largest num_rendered seen (x headroom)."""
        worst = 0
        for v in views_dev:
            self.render(v)
            worst = max(worst, self.status()[0])
        if worst * self.headroom > self.capacity:
            self._reserve(int(worst * self.headroom) + 1024)
        return worst


class ViewBatch:
    """All cameras of a render call set up on the device by ONE kernel: what the reference's caller does
    per view on the host in get_rasterize_param_from_camera (rigid inverse, transposes, a bmm, four small H2D copies;
    simple_raw_render.py:79-112).  `batch[k]` is the view tuple FrameRenderer.enqueue takes."""

    def __init__(self, c2w, fov_deg: float, device):
        import math
        self.dev = torch.device(device)
        c2w = torch.as_tensor(c2w, dtype=torch.float32).reshape(-1, 4, 4)
        self.c2w = c2w.to(self.dev, non_blocking=True)  # one upload for all views
        self.buf = _C.make_views(self.c2w, fov_deg, fov_deg)
        self.tanfov = math.tan(fov_deg / 180.0 * math.pi)  # the reference passes the FULL angle (simple_raw_render.py:102)

    def __len__(self) -> int:
        return int(self.buf.shape[0])

    def __getitem__(self, k: int):
        row = self.buf[k]
        return (row[0:16], row[16:32], row[32:35], self.tanfov, self.tanfov)


def render_passes(fr, views: "ViewBatch", normals: Optional[torch.Tensor] = None) -> dict:
    """The four raster passes of the reference caller's `render()` for every view of a batch
    (simple_raw_render.py:411-522): world position (`xyz_w`, colours = the Gaussian centres), RGB from the SH
    coefficients (`rgb`), hit map (`hitmap`, colours = 1) and -- if `normals` (P,3) is given -- the camera-facing
    normal map (`normal`).  One preprocess / sort / binning and ONE list walk per view instead of four rasterizer
    calls; with `downsample == 2` the bilinear x0.5 of `_rasterize` (:281-284) happens in the blend epilogue.
    `fr` is a FrameRenderer (views one after the other on the current stream) or a FramePipeline (views round-robin
    on its lanes).  Returns (N, h, w, 3) tensors (permuted views of (N,3,h,w) storage, as the reference returns
    them), bit-identical to the reference call sequence; valid once the current stream has caught up.

    The normal pass reproduces `normalize_camera_normal` including its quirk: the flipped normals of view j are the
    input of view j+1 (`colors_precomp_i` is reassigned inside the view loop, :264-268)."""
    pipe = fr if isinstance(fr, FramePipeline) else None
    first = pipe.lanes[0] if pipe is not None else fr
    N = len(views)
    h, w = first.H // first.downsample, first.W // first.downsample
    dev = first.dev
    names = ["rgb", "xyz_w", "hitmap"] + (["normal"] if normals is not None else [])
    out = {n: torch.empty((N, 3, h, w), dtype=torch.float32, device=dev) for n in names}
    ones = torch.ones_like(first.means3D)
    per_view_normals = []
    if normals is not None:  # the flips chain from view to view: all of them first, on the current stream
        nrm = normals.to(dev, torch.float32).contiguous()
        for k in range(N):
            camera_dir = first.means3D - views[k][2].reshape(1, 3)               # means3D - camera origin
            sgn = (torch.sum(camera_dir * nrm, -1, keepdim=True) > 0).float() * 2 - 1
            nrm = nrm * (-1) * sgn
            per_view_normals.append(nrm)
    if pipe is not None:
        pipe.begin()
    for k in range(N):
        extra = [(first.means3D, out["xyz_w"][k]), (ones, out["hitmap"][k])]
        if normals is not None:
            extra.append((per_view_normals[k], out["normal"][k]))
        if pipe is None:
            fr.enqueue(views[k], out_color=out["rgb"][k], extra_passes=extra, slot=k)
        else:
            lane = pipe.count % len(pipe.lanes)
            pipe.count += 1
            with torch.cuda.stream(pipe.streams[lane]):
                pipe.lanes[lane].enqueue(views[k], out_color=out["rgb"][k], extra_passes=extra, slot=k)
    if pipe is not None:
        pipe.end()
    first._keep_passes = (per_view_normals, ones)  # alive until the streams have consumed them
    return {n: t.permute(0, 2, 3, 1) for n, t in out.items()}


def host_block(cloud: dict, pin: bool = True) -> dict:
    """The Gaussian attributes of a cloud as ONE host buffer: returns the cloud dict with its five attribute arrays
    replaced by views into a single (pinned) float32 block, arrays back to back (each start a multiple of 64 floats),
    plus the block under "_flat" and the (name, offset, shape) layout under "_layout".
    FramePipeline.enqueue_host then moves the whole cloud with one host->device copy instead of five (measured on the
    B200 box: 50.1 instead of 45.6 GB/s with the image download running the other way, tools/pcie_probe.py)."""
    names = FramePipeline._ATTRS
    layout, off = [], 0
    for n in names:
        t = cloud[n]
        layout.append((n, off, tuple(t.shape)))
        off += -(-t.numel() // 64) * 64
    flat = torch.empty(off, dtype=torch.float32)
    if pin:
        flat = flat.pin_memory()
    out = dict(cloud)
    for n, o, shp in layout:
        cnt = 1
        for d in shp:
            cnt *= d
        v = flat[o:o + cnt].view(shp)
        v.copy_(cloud[n].to(torch.float32))
        out[n] = v
    out["_flat"], out["_layout"] = flat, layout
    return out


class FramePipeline:
    """Several frames in flight on separate CUDA streams, each with its own workspaces (the cloud is shared).

    A frame of a THuman-shaped cloud ends with a long, thinly populated tail (a few silhouette pixel blocks walk lists
    that are an order of magnitude longer than the average) and starts with a chain of short, latency-bound binning
    kernels; neither fills the GPU.  Frames of an orbit are independent, so rendering them round-robin on `depth`
    streams lets the tail of one frame overlap the busy phases of the next: throughput goes up, per-frame latency is
    unchanged, and every frame is still bit-identical to a frame rendered alone.
    """

    def __init__(self, cloud: dict, width: int, height: int, bg, device, depth: int = 3, capacity: int = 0,
                 headroom: float = 1.3, downsample: int = 1):
        self.lanes = []
        for k in range(max(1, int(depth))):
            self.lanes.append(FrameRenderer(cloud, width, height, bg, device, capacity=capacity, headroom=headroom,
                                            share=self.lanes[0] if self.lanes else None, downsample=downsample,
                                            team_after=-1 if int(depth) > 1 else 0))  # frames in flight: throughput mode
        self.dev = self.lanes[0].dev
        with torch.cuda.device(self.dev):
            self.streams = [torch.cuda.Stream(self.dev) for _ in self.lanes]
        self.count = 0

    @property
    def depth(self) -> int:
        return len(self.lanes)

    def upload_view(self, view):
        return self.lanes[0].upload_view(view)

    def calibrate(self, views_dev) -> int:
        worst = self.lanes[0].calibrate(views_dev)
        for ln in self.lanes[1:]:
            if ln.capacity < self.lanes[0].capacity:
                ln._reserve(self.lanes[0].capacity)
        return worst

    def begin(self) -> None:
        """Orders the lanes behind the work already queued on the current stream."""
        ev = torch.cuda.current_stream(self.dev).record_event()
        for st in self.streams + ([self._feed] if hasattr(self, "_feed") else []):
            st.wait_event(ev)

    def enqueue(self, view_dev, slot: int = 0, tile_rows=None):
        """Queues one frame on the next lane; returns (lane index, colour tensor of that lane)."""
        k = self.count % len(self.lanes)
        self.count += 1
        with torch.cuda.stream(self.streams[k]):
            out = self.lanes[k].enqueue(view_dev, tile_rows=tile_rows, slot=slot)
        return k, out

    def capture_graphs(self, tanfov: Tuple[float, float]) -> None:
        """One CUDA graph per lane (FrameRenderer.capture_graph)."""
        for ln, st in zip(self.lanes, self.streams):
            with torch.cuda.stream(st):
                ln.capture_graph(tanfov)
        torch.cuda.synchronize(self.dev)

    def enqueue_graph(self, view_row: torch.Tensor):
        """Queues one frame on the next lane through its captured graph; returns (lane index, colour tensor)."""
        k = self.count % len(self.lanes)
        self.count += 1
        with torch.cuda.stream(self.streams[k]):
            out = self.lanes[k].enqueue_graph(view_row)
        return k, out

    _ATTRS = ("means3D", "opacities", "scales", "rotations", "shs")

    def _own_inputs(self, ln, rows: int, host_cloud: dict) -> None:
        """Private device copies of the Gaussian attributes for one lane, shaped like the host arrays that will be copied
        into them (`rows` >= P rows each; the renderer reads the first P)."""
        key = (rows,) + tuple(tuple(host_cloud[n].shape[1:]) for n in self._ATTRS)
        if getattr(ln, "_in_key", None) == key:
            return
        ln._in = {}
        for n in self._ATTRS:
            ln._in[n] = torch.empty((rows,) + tuple(host_cloud[n].shape[1:]), dtype=torch.float32, device=self.dev)
            setattr(ln, n, ln._in[n][: ln.P])
        ln._in_key = key
        ln._flat_key = None
        ln._view_dev = (torch.empty(4, 4, device=self.dev), torch.empty(4, 4, device=self.dev),
                        torch.empty(3, device=self.dev))
        ln._in_rows = rows
        ln._consumed = None  # event: the lane's last frame has read its inputs

    def _own_inputs_flat(self, ln, layout, numel: int, world: int) -> int:
        """Private device block of one lane for a host_block cloud: the lane's attribute tensors become views into it
        (same offsets as the host block).  Padded to a multiple of `world` equal slices; returns the slice length."""
        S = -(-numel // (64 * world)) * 64
        if getattr(ln, "_flat_key", None) != (numel, world):
            ln._flat = torch.empty(S * world, dtype=torch.float32, device=self.dev)
            for n, o, shp in layout:
                cnt = 1
                for d in shp:
                    cnt *= d
                setattr(ln, n, ln._flat[o:o + cnt].view(shp))
            ln._view_dev = (torch.empty(4, 4, device=self.dev), torch.empty(4, 4, device=self.dev),
                            torch.empty(3, device=self.dev))
            ln._flat_key, ln._in_key, ln._consumed = (numel, world), None, None
        return S

    def enqueue_host(self, host_cloud: dict, host_view, tanfov, out_host: torch.Tensor, slot: int = 0,
                     group=None) -> int:
        """One frame whose inputs live in (pinned) HOST memory: uploads the Gaussian attributes and the camera of
        this frame on the lane's stream, renders, and downloads the image into `out_host` (pinned, (3,H,W)).
        Copies of one frame overlap the kernels of the frames on the other lanes.  Returns the lane index; the
        image is valid once the lane's stream (or `end()` + the current stream) has been synchronised.

        group (a torch.distributed process group of the GPUs of one node that all render views of the SAME cloud in
        lock step -- the view-parallel orbit): the cloud crosses PCIe once per step and node instead of
        once per rank: every rank uploads rows [r S, (r+1) S) of each attribute array (S = ceil(P / world)) and one
        all-gather per array over NVLink completes the copy on every GPU."""
        k = self.count % len(self.lanes)
        self.count += 1
        ln = self.lanes[k]
        flat = host_cloud.get("_flat")
        if flat is not None:  # the cloud as one host block (host_block): one copy (per rank: one slice + one all-gather)
            import torch.distributed as dist
            world = dist.get_world_size(group) if group is not None else 1
            rank = dist.get_rank(group) if group is not None else 0
            S = self._own_inputs_flat(ln, host_cloud["_layout"], flat.numel(), world)
            if group is None:
                with torch.cuda.stream(self.streams[k]):
                    ln._flat[: flat.numel()].copy_(flat, non_blocking=True)
                    for dst, src in zip(ln._view_dev, host_view):
                        dst.copy_(src, non_blocking=True)
                    out = ln.enqueue(ln._view_dev + (tanfov[0], tanfov[1]), slot=slot)
                    out_host.copy_(out, non_blocking=True)
                return k
            if not hasattr(self, "_feed"):
                self._feed = torch.cuda.Stream(self.dev)
            a, b = min(flat.numel(), rank * S), min(flat.numel(), (rank + 1) * S)
            with torch.cuda.stream(self._feed):
                if ln._consumed is not None:
                    self._feed.wait_event(ln._consumed)
                if b > a:
                    ln._flat[a:b].copy_(flat[a:b], non_blocking=True)
                dist.all_gather_into_tensor(ln._flat, ln._flat[rank * S:(rank + 1) * S], group=group)
                fed = self._feed.record_event()
            with torch.cuda.stream(self.streams[k]):
                self.streams[k].wait_event(fed)
                for dst, src in zip(ln._view_dev, host_view):
                    dst.copy_(src, non_blocking=True)
                out = ln.enqueue(ln._view_dev + (tanfov[0], tanfov[1]), slot=slot)
                ln._consumed = self.streams[k].record_event()
                out_host.copy_(out, non_blocking=True)
            return k
        if group is None:
            self._own_inputs(ln, ln.P, host_cloud)
            with torch.cuda.stream(self.streams[k]):
                for n in self._ATTRS:
                    getattr(ln, n).copy_(host_cloud[n], non_blocking=True)
                for dst, src in zip(ln._view_dev, host_view):
                    dst.copy_(src, non_blocking=True)
                out = ln.enqueue(ln._view_dev + (tanfov[0], tanfov[1]), slot=slot)
                out_host.copy_(out, non_blocking=True)
            return k
        import torch.distributed as dist
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        S = (ln.P + world - 1) // world
        self._own_inputs(ln, S * world, host_cloud)
        if not hasattr(self, "_feed"):  # all collectives of this pipeline are issued from ONE stream, in step order
            self._feed = torch.cuda.Stream(self.dev)
        a, b = min(ln.P, rank * S), min(ln.P, (rank + 1) * S)
        with torch.cuda.stream(self._feed):
            if ln._consumed is not None:
                self._feed.wait_event(ln._consumed)  # frame i - depth of this lane has read the old inputs
            for n in self._ATTRS:
                buf = ln._in[n]
                if b > a:
                    buf[a:b].copy_(host_cloud[n][a:b], non_blocking=True)
                dist.all_gather_into_tensor(buf, buf[rank * S:(rank + 1) * S], group=group)
            fed = self._feed.record_event()
        with torch.cuda.stream(self.streams[k]):
            self.streams[k].wait_event(fed)
            for dst, src in zip(ln._view_dev, host_view):
                dst.copy_(src, non_blocking=True)
            out = ln.enqueue(ln._view_dev + (tanfov[0], tanfov[1]), slot=slot)
            ln._consumed = self.streams[k].record_event()
            out_host.copy_(out, non_blocking=True)
        return k

    def end(self) -> None:
        """Orders the current stream behind every lane (call before recording the closing event)."""
        cur = torch.cuda.current_stream(self.dev)
        for st in self.streams:
            cur.wait_stream(st)
