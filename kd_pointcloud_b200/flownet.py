"""Bi-PointFlowNet (``PointConvBidirection``) on the kdpc layers.

Mirror of the reference's teacher ``models_bid_pointconv.PointConvBidirection``
(models_bid_pointconv.py:14-207) and student ``models_bid_lighttoken_res.PointConvBidirection``
(models_bid_lighttoken_res.py:14-189) — the two are the same architecture with identical
``state_dict`` keys; ``weightnet`` is the student's only knob.  Same attribute names, so
reference checkpoints load; same ``forward(xyz1, xyz2, color1, color2)`` 8-tuple.

What differs is the schedule, not the math:
  * activations stay point-major end to end (the reference permutes/copies around every op);
  * both clouds go through the weight-shared encoder as ONE batch of 2B clouds (one FPS launch
    covers 2B CTAs, half the launches everywhere); the encoder has no BatchNorm, so this is
    exactly equivalent;
  * each 3-NN / kNN index set is computed once and shared by every consumer of the same
    (query, candidate) pair.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import functional as KF
from .functional import cm, knn_idx
from .pointconv_util import (Conv1d, CrossLayerLight, PointConvD, PointWarping, SceneFlowEstimatorResidual,
                             UpsampleFlow)

scale = 1.0
OVERLAP_SAMPLING = False     # opt-in: measured 11 % SLOWER on B200 - the cluster FPS wants all its SMs at once and the
                             # persistent tcgen05 kernels of the main stream hold every SM (results are identical);
                             # with those grids capped to the 84 SMs the FPS clusters leave free it is 15 % slower:
                             # the co-resident kNN warps stretch the FPS latency chain
_SIDE_STREAMS = {}


class _SamplePyramid:
    """FPS + gather for all four levels, issued on a side CUDA stream (fork/join that a CUDA graph capture records
    as parallel branches).  ``level(i)`` makes the current stream wait for level i and hands out (fps_idx, new_xyz)."""

    def __init__(self, pc_l0: torch.Tensor, npoints):
        self.overlapped = bool(OVERLAP_SAMPLING and pc_l0.is_cuda)
        self.items, self.events = [], []
        if not self.overlapped:
            xyz = pc_l0
            for n in npoints:
                idx = KF.furthest_point_sample(xyz, n)
                xyz = KF.gather_rows(xyz, idx)
                self.items.append((idx, xyz))
            return
        dev = pc_l0.device
        main = torch.cuda.current_stream(dev)
        side = _SIDE_STREAMS.get(dev)
        if side is None:
            side = _SIDE_STREAMS[dev] = torch.cuda.Stream(device=dev)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            xyz = pc_l0
            for n in npoints:
                idx = KF.furthest_point_sample(xyz, n)
                xyz = KF.gather_rows(xyz, idx)
                ev = torch.cuda.Event()
                ev.record(side)
                self.items.append((idx, xyz))
                self.events.append(ev)
        pc_l0.record_stream(side)
        self._main = main

    def level(self, i: int):
        idx, xyz = self.items[i]
        if self.overlapped:
            torch.cuda.current_stream(idx.device).wait_event(self.events[i])
            idx.record_stream(torch.cuda.current_stream(idx.device))
            xyz.record_stream(torch.cuda.current_stream(idx.device))
        return idx, xyz


# The 2B-cloud batch and its sampling pyramid depend on the input coordinates only.  Teacher and student of a KD step
# (training.kd_step) see the SAME clouds, so the second model reuses the first one's tensors - and with them every
# coordinate-only kNN / spatial sort / CSR downstream, whose caches are keyed by tensor identity.  (The reference runs
# FPS and all 45 kNNs in both models.)  Dropped with the other per-batch caches by functional.clear_caches().
_GEOMETRY_CACHE = KF._LRU(4)
KF.register_batch_cache(_GEOMETRY_CACHE)


class PointConvBidirection(nn.Module):
    def __init__(self, weightnet: int = 16):
        super().__init__()
        flow_nei, feat_nei = 32, 16
        self.scale = scale
        # l0: 8192
        self.level0 = Conv1d(3, 32)
        self.level0_1 = Conv1d(32, 32)
        self.cross0 = CrossLayerLight(flow_nei, 32 + 32, [32, 32], [32, 32])
        self.flow0 = SceneFlowEstimatorResidual(32 + 64, 32, weightnet=weightnet)
        self.level0_2 = Conv1d(32, 64)
        # l1: 2048
        self.level1 = PointConvD(2048, feat_nei, 64 + 3, 64, weightnet=weightnet)
        self.cross1 = CrossLayerLight(flow_nei, 64 + 32, [64, 64], [64, 64])
        self.flow1 = SceneFlowEstimatorResidual(64 + 64, 64, weightnet=weightnet)
        self.level1_0 = Conv1d(64, 64)
        self.level1_1 = Conv1d(64, 128)
        # l2: 512
        self.level2 = PointConvD(512, feat_nei, 128 + 3, 128, weightnet=weightnet)
        self.cross2 = CrossLayerLight(flow_nei, 128 + 64, [128, 128], [128, 128])
        self.flow2 = SceneFlowEstimatorResidual(128 + 64, 128, weightnet=weightnet)
        self.level2_0 = Conv1d(128, 128)
        self.level2_1 = Conv1d(128, 256)
        # l3: 256
        self.level3 = PointConvD(256, feat_nei, 256 + 3, 256, weightnet=weightnet)
        self.cross3 = CrossLayerLight(flow_nei, 256 + 64, [256, 256], [256, 256])
        self.flow3 = SceneFlowEstimatorResidual(256, 256, weightnet=weightnet)
        self.level3_0 = Conv1d(256, 256)
        self.level3_1 = Conv1d(256, 512)
        # l4: 64
        self.level4 = PointConvD(64, feat_nei, 512 + 3, 256, weightnet=weightnet)
        # deconv
        self.deconv4_3 = Conv1d(256, 64)
        self.deconv3_2 = Conv1d(256, 64)
        self.deconv2_1 = Conv1d(128, 32)
        self.deconv1_0 = Conv1d(64, 32)
        self.warping = PointWarping()
        self.upsample = UpsampleFlow()

    def _sample_pyramid(self, pc_l0):
        return _SamplePyramid(pc_l0, [self.level1.npoint, self.level2.npoint, self.level3.npoint, self.level4.npoint])

    def sample_geometry(self, xyz1, xyz2):
        """The part of the forward that depends on the input COORDINATES only: the 2B-cloud batch and its four-level
        FPS pyramid.  ``forward(..., geometry=g)`` consumes it; the runner computes it for batch i+1 on a second stream
        while batch i's forward runs (the pyramid is a 1.3 ms latency chain that leaves most SMs idle)."""
        pc_l0 = torch.cat([xyz1, xyz2], dim=0).contiguous()
        return pc_l0, self._sample_pyramid(pc_l0)

    def precompute_neighbours(self, geometry) -> None:
        """Every neighbour search of the forward that depends on the INPUT coordinates only (not on a warped cloud):
        the PointConvD groupings, the four dense<-sparse 3-NN sets, the flow estimators' self-kNN and the level-3 cost
        volume.  Issued here (same calls, same tensors), they land in functional's kNN / sort caches; a forward that
        finds them there launches only the three warped-cloud searches.  The runner does this for batch i+1 beside the
        forward of batch i."""
        pc_l0, pyramid = geometry
        B = pc_l0.shape[0] // 2
        pcs = [pc_l0] + [pyramid.level(i)[1] for i in range(4)]
        for lvl, layer in enumerate((self.level1, self.level2, self.level3, self.level4)):
            knn_idx(layer.nsample, pcs[lvl], pcs[lvl + 1])                       # PointConvD: queries = sampled points
        knn_idx(3, pcs[4], pcs[3])                                               # up(pc_l3, pc_l4, f_l4)
        for lvl in (2, 1, 0):
            knn_idx(3, pcs[lvl + 1], pcs[lvl])                                   # up32 / up21 / up10
        for lvl, est in ((3, self.flow3), (2, self.flow2), (1, self.flow1), (0, self.flow0)):
            p1 = pcs[lvl][:B]
            knn_idx(est.pointconv_list[0].nsample, p1, p1)                       # SceneFlowEstimatorResidual: self-kNN
        p1, p2 = pcs[3][:B], pcs[3][B:]
        knn_idx(self.cross3.nsample, p2, p1)                                     # cross3 (level 3 is not warped)
        knn_idx(self.cross3.nsample, p1, p2)

    def forward(self, xyz1, xyz2, color1, color2, geometry=None):
        # xyz*, color*: [B,N,3]   (models_bid_pointconv.py:74-92)
        B = xyz1.shape[0]
        up = self.upsample.forward_pm
        cat_c = lambda *ts: torch.cat(ts, dim=2)
        both = KF.joined                      # cat along the batch axis (no copy when the halves already share a tensor)
        # Inference: the per-level concatenations [encoder features | upsampled decoder features] are never copied
        # together - the two 1x1 convolutions that produce the parts write straight into the column blocks of one
        # buffer (the kernels take row strides), and the layers that read a part read the view.  Same values.
        free = KF.concat_free(xyz1, xyz2, color1, color2)

        def parts(rows_like, c_first, c_second):
            """(whole, first, second): a [2B,N,c_first+c_second] buffer and its two column blocks, or Nones."""
            if not free:
                return None, None, None
            whole = rows_like.new_empty(tuple(rows_like.shape[:2]) + (c_first + c_second,))
            return whole, whole[:, :, :c_first], whole[:, :, c_first:]

        # ---- encoder: clouds 1 and 2 as one batch of 2B (weights are shared, no BN) --------------
        npts = (self.level1.npoint, self.level2.npoint, self.level3.npoint, self.level4.npoint)
        gkey = (KF._tkey(xyz1), KF._tkey(xyz2), npts)
        hit = _GEOMETRY_CACHE.get(gkey) if (KF._CACHE_ENABLED and not xyz1.requires_grad and not xyz2.requires_grad) else None
        if geometry is not None:
            pc_l0, pyramid = geometry
        elif hit is not None:
            pc_l0, pyramid = hit[0], hit[1]
        else:
            pc_l0 = both(xyz1, xyz2).contiguous()
            # The sampling pyramid depends on coordinates only and is a chain of latency-bound kernels on <= 64 SMs
            # (optionally on a side stream while this stream does the level-0 convolutions and the level-0 self-kNN).
            pyramid = self._sample_pyramid(pc_l0)
            if KF._CACHE_ENABLED and not pyramid.overlapped:
                _GEOMETRY_CACHE.put(gkey, (pc_l0, pyramid, xyz1, xyz2))
        c_feat_l0, c0a, c0b = parts(pc_l0, self.level0_1.out_channels, self.deconv1_0.out_channels)
        f_l0 = self.level0_1.forward_pm(self.level0.forward_pm(torch.cat([color1, color2], dim=0)), out=c0a)
        f_l0_1 = self.level0_2.forward_pm(f_l0)
        if pyramid.overlapped:
            knn_idx(self.flow0.pointconv_list[0].nsample, pc_l0[:B], pc_l0[:B])      # flow0's neighbourhoods (cached)

        pc_l1, f_l1, fps_l1 = self.level1.forward_pm(pc_l0, f_l0_1, pyramid.level(0))
        c_feat_l1, c1a, c1b = parts(pc_l1, self.level1_0.out_channels, self.deconv2_1.out_channels)
        f_l1 = self.level1_0.forward_pm(f_l1, out=c1a)
        f_l1_2 = self.level1_1.forward_pm(f_l1)

        pc_l2, f_l2, fps_l2 = self.level2.forward_pm(pc_l1, f_l1_2, pyramid.level(1))
        c_feat_l2, c2a, c2b = parts(pc_l2, self.level2_0.out_channels, self.deconv3_2.out_channels)
        f_l2 = self.level2_0.forward_pm(f_l2, out=c2a)
        f_l2_3 = self.level2_1.forward_pm(f_l2)

        pc_l3, f_l3, fps_l3 = self.level3.forward_pm(pc_l2, f_l2_3, pyramid.level(2))
        c_feat_l3, c3a, c3b = parts(pc_l3, self.level3_0.out_channels, self.deconv4_3.out_channels)
        f_l3 = self.level3_0.forward_pm(f_l3, out=c3a)
        f_l3_4 = self.level3_1.forward_pm(f_l3)

        pc_l4, f_l4, _ = self.level4.forward_pm(pc_l3, f_l3_4, pyramid.level(3))
        f_l4_3 = self.deconv4_3.forward_pm(up(pc_l3, pc_l4, f_l4), out=c3b)

        # 3-NN index sets dense<-sparse, computed once for both clouds and reused below
        up32 = knn_idx(3, pc_l3, pc_l2)
        up21 = knn_idx(3, pc_l2, pc_l1)
        up10 = knn_idx(3, pc_l1, pc_l0)

        h = lambda t: (t[:B], t[B:])
        (pc1_l0, pc2_l0), (pc1_l1, pc2_l1), (pc1_l2, pc2_l2), (pc1_l3, pc2_l3) = h(pc_l0), h(pc_l1), h(pc_l2), h(pc_l3)
        feat1_l0, feat2_l0 = h(f_l0)
        feat1_l1, feat2_l1 = h(f_l1)
        feat1_l2, feat2_l2 = h(f_l2)
        feat1_l3, feat2_l3 = h(f_l3)

        # ---- l3 -------------------------------------------------------------------------------
        if not free:
            c_feat_l3 = cat_c(f_l3, f_l4_3)
        c_feat1_l3, c_feat2_l3 = h(c_feat_l3)
        feat1_new_l3, feat2_new_l3, cross3 = self.cross3.forward_pm(pc1_l3, pc2_l3, c_feat1_l3, c_feat2_l3, c_feat_l3)
        feat3, flow3 = self.flow3.forward_pm(pc1_l3, feat1_l3, cross3)

        f_l3_2 = self.deconv3_2.forward_pm(up(pc_l2, pc_l3, both(feat1_new_l3, feat2_new_l3), idx=up32), out=c2b)
        if not free:
            c_feat_l2 = cat_c(f_l2, f_l3_2)
        c_feat1_l2, c_feat2_l2 = h(c_feat_l2)

        # ---- l2 -------------------------------------------------------------------------------
        up_flow2 = up(pc1_l2, pc1_l3, self.scale * flow3, idx=up32[:B])
        pc2_l2_warp = self.warping.forward_pm(pc1_l2, pc2_l2, up_flow2)
        feat1_new_l2, feat2_new_l2, cross2 = self.cross2.forward_pm(pc1_l2, pc2_l2_warp, c_feat1_l2, c_feat2_l2, c_feat_l2)
        feat3_up = up(pc1_l2, pc1_l3, feat3, idx=up32[:B])
        feat2, flow2 = self.flow2.forward_pm(pc1_l2, (feat1_l2, feat3_up), cross2, up_flow2)

        f_l2_1 = self.deconv2_1.forward_pm(up(pc_l1, pc_l2, both(feat1_new_l2, feat2_new_l2), idx=up21), out=c1b)
        if not free:
            c_feat_l1 = cat_c(f_l1, f_l2_1)
        c_feat1_l1, c_feat2_l1 = h(c_feat_l1)

        # ---- l1 -------------------------------------------------------------------------------
        up_flow1 = up(pc1_l1, pc1_l2, self.scale * flow2, idx=up21[:B])
        pc2_l1_warp = self.warping.forward_pm(pc1_l1, pc2_l1, up_flow1)
        feat1_new_l1, feat2_new_l1, cross1 = self.cross1.forward_pm(pc1_l1, pc2_l1_warp, c_feat1_l1, c_feat2_l1, c_feat_l1)
        feat2_up = up(pc1_l1, pc1_l2, feat2, idx=up21[:B])
        feat1, flow1 = self.flow1.forward_pm(pc1_l1, (feat1_l1, feat2_up), cross1, up_flow1)

        f_l1_0 = self.deconv1_0.forward_pm(up(pc_l0, pc_l1, both(feat1_new_l1, feat2_new_l1), idx=up10), out=c0b)
        if not free:
            c_feat_l0 = cat_c(f_l0, f_l1_0)
        c_feat1_l0, c_feat2_l0 = h(c_feat_l0)

        # ---- l0 -------------------------------------------------------------------------------
        up_flow0 = up(pc1_l0, pc1_l1, self.scale * flow1, idx=up10[:B])
        pc2_l0_warp = self.warping.forward_pm(pc1_l0, pc2_l0, up_flow0)
        _, _, cross0 = self.cross0.forward_pm(pc1_l0, pc2_l0_warp, c_feat1_l0, c_feat2_l0, c_feat_l0)
        feat1_up = up(pc1_l0, pc1_l1, feat1, idx=up10[:B])
        _, flow0 = self.flow0.forward_pm(pc1_l0, (feat1_l0, feat1_up), cross0, up_flow0)

        # ---- outputs in the reference's [B,C,N] layout (models_bid_pointconv.py:198-207) ------
        flows = [cm(flow0), cm(flow1), cm(flow2), cm(flow3)]
        pc1 = [cm(pc1_l0), cm(pc1_l1), cm(pc1_l2), cm(pc1_l3)]
        pc2 = [cm(pc2_l0), cm(pc2_l1), cm(pc2_l2), cm(pc2_l3)]
        fps_pc1_idxs = [fps_l1[:B], fps_l2[:B], fps_l3[:B]]
        fps_pc2_idxs = [fps_l1[B:], fps_l2[B:], fps_l3[B:]]
        s = lambda t, i: cm(t[:B] if i == 0 else t[B:])
        feat1s = [s(t, 0) for t in (f_l0_1, f_l1_2, f_l2_3, f_l3_4, f_l3_2, f_l2_1, f_l1_0)]
        feat2s = [s(t, 1) for t in (f_l0_1, f_l1_2, f_l2_3, f_l3_4, f_l3_2, f_l2_1, f_l1_0)]
        crosses = [cm(cross0), cm(cross1), cm(cross2), cm(cross3)]
        return flows, fps_pc1_idxs, fps_pc2_idxs, pc1, pc2, feat1s, feat2s, crosses


def teacher() -> PointConvBidirection:
    """models_bid_pointconv.PointConvBidirection()"""
    return PointConvBidirection(weightnet=16)


def student(weightnet: int = 16) -> PointConvBidirection:
    """models_bid_lighttoken_res.PointConvBidirection() (weightnet = 16, models_bid_lighttoken_res.py:20)"""
    return PointConvBidirection(weightnet=weightnet)
