"""The input side of Network.forward on the device: voxelisation and the two sparse-conv
towers, inference only.

    Backbone_SPCONV(dims, stride_layers, cfg)      models/Modules.py:100-159 — same module / parameter names
                                                   (module{1..4}.{0,1}.layers.0.weight (3,3,3,Cin,Cout), layers.1 = BN1d),
                                                   so a reference checkpoint's towers load
    SparseTowers(backbone_inp, backbone_tmp, ...)  both towers of a batch: voxel sets, rulebooks, 8 convolutions and 4
                                                   average pools per tower, ~20 launches for the pair, no host sync
    voxelization_idx / voxelization                libs/pointgroup_ops (op-level drop-ins)

What replaces what: the dataloader's CPU hash map (voxelize.cpp:57-163) and spconv's dense-grid index generation
(indice.cu.h) become bit-grid kernels (csrc/sparse_index.cu); spconv's per-offset gather -> cuBLAS mm -> scatter-add
(spconv_ops.h:253-349) becomes one output-stationary tcgen05 kernel per layer (csrc/sparse_conv.cu) with eval-mode
BatchNorm1d folded into the weights and the ReLU in the epilogue.  Activations between the layers of a tower are fp16
operand rows (the format of the whole inference path); the four pyramid levels the point-feature interpolation reads
are fp32.  Rows of every sparse tensor are in the reference's order (batch, then linear voxel index: spconv_ops.h:122
sorts them), packed densely over the batch in buffers of fixed capacity; unused rows carry batch id == B.
"""
import ctypes
import math
import types

import numpy as np
import torch
import torch.nn as nn

from . import _lib as L

NSETS, NOPS = 9, 12
# bench.py sets this to a list to collect (start event, end event, rulebook op, c_in, c_out) around every sparse-conv launch
CONV_EVENTS = None
DIMS = (7, 16, 32, 32, 64, 64, 128, 128, 256)


class _SparseConvParams(nn.Module):
    """Parameter holder with spconv's layout and initialisation (libs/spconv/spconv/conv.py:98-111)."""

    def __init__(self, in_channels, out_channels, subm):
        super().__init__()
        self.in_channels, self.out_channels, self.subm = in_channels, out_channels, subm
        self.weight = nn.Parameter(torch.empty(3, 3, 3, in_channels, out_channels))
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))

    def forward(self, x):
        raise RuntimeError("sparse convolutions run inside backbone.SparseTowers (whole towers on the device); "
                           "dcl_net_b200 has no per-module or CPU path for them")


class BasicBlock_SPCONV(nn.Module):
    """models/Modules.py:11-56: conv -> BatchNorm1d -> ReLU under `layers` (bias=False, norm=True, act='relu')."""

    def __init__(self, subm, dim_in, dim_out, **_unused):
        super().__init__()
        self.layers = nn.Sequential(_SparseConvParams(dim_in, dim_out, subm), nn.BatchNorm1d(dim_out), nn.ReLU())


class Backbone_SPCONV(nn.Module):
    def __init__(self, dims=DIMS, stride_layers=(1, 3, 5), cfg=None, norm=True):
        super().__init__()
        assert tuple(dims) == DIMS and tuple(stride_layers) == (1, 3, 5) and norm, \
            "the device towers implement the configuration of models/DCL_Net.py:47-52"
        modules = [[] for _ in range(len(stride_layers) + 1)]
        mi = 0
        for i in range(len(dims) - 1):
            subm = not ((i - 1) in stride_layers or i == 0)
            modules[mi].append(BasicBlock_SPCONV(subm, dims[i], dims[i + 1]))
            if i in stride_layers:
                mi += 1
        self.module1, self.module2, self.module3, self.module4 = (nn.Sequential(*m) for m in modules)

    def blocks(self):
        return [blk for mod in (self.module1, self.module2, self.module3, self.module4) for blk in mod]

    def forward(self, inputs):
        raise RuntimeError("Backbone_SPCONV runs inside backbone.SparseTowers (both towers, whole batch)")


def pack_conv_weight(w, scale, chan_map=None, cin_pad=None):
    """(3,3,3,Cin,Cout) fp32 weights with the BatchNorm scale folded -> the packed fp16 hi/lo image of
    dcl_spb_conv3: reduction axis kv = k*cin_pad + c padded to a multiple of 64; per stage of 64 kv: [hi | lo] images of Cout rows x 64, K-major 8x8 core matrices:
        byte(o, kk) = (o/8)*1024 + (kk/8)*128 + (o%8)*16 + (kk%8)*2.
    chan_map[c] = source input channel of operand channel c (-1: zero), for the first layer's [value | remainder] rows."""
    cin, cout = w.shape[3], w.shape[4]
    cin_pad = cin_pad or cin
    wf = (w.reshape(27, cin, cout) * scale.view(1, 1, cout)).float()
    if chan_map is None:
        chan_map = list(range(cin)) + [-1] * (cin_pad - cin)
    src = torch.as_tensor([max(c, 0) for c in chan_map], dtype=torch.long, device=wf.device)
    live = torch.as_tensor([float(c >= 0) for c in chan_map], dtype=wf.dtype, device=wf.device)
    wv = wf.index_select(1, src) * live.view(1, -1, 1)
    kp = (27 * cin_pad + 63) // 64 * 64
    flat = wf.new_zeros(kp, cout)
    flat[:27 * cin_pad] = wv.reshape(27 * cin_pad, cout)
    nt = cout                                                    # one n-tile (NT = Cout <= 256)
    wt = flat.t().contiguous()                                   # (cout, kp)
    hi = wt.to(torch.float16)
    lo = (wt - hi.float()).to(torch.float16)

    def img(x):   # (tile, og, o8, s, kc8, k8) -> (tile, s, og, kc8, o8, k8)
        return x.view(cout // nt, nt // 8, 8, kp // 64, 8, 8).permute(0, 3, 1, 4, 2, 5)
    packed = torch.stack([img(hi), img(lo)], dim=2).contiguous()  # (tile, s, half, og, kc8, o8, k8)
    return packed.view(torch.uint8).reshape(-1)


def _bn_affine(bn):
    s = bn.weight / torch.sqrt(bn.running_var + bn.eps)
    return s, bn.bias - bn.running_mean * s


# operand channels of the first convolution: [1, rgb, xyz] then the fp16 remainders of rgb, xyz (sparse_index.cu)
_CONV0_CHANNELS = [0, 1, 2, 3, 4, 5, 6, 1, 2, 3, 4, 5, 6, -1, -1, -1]


class _TowerWeights:
    def __init__(self, backbone):
        self.layers = []   # per level: ((w_conv, shift_conv), (w_subm, shift_subm))
        with torch.no_grad():
            blocks = backbone.blocks()
            for level in range(4):
                pair = []
                for j, blk in enumerate(blocks[2 * level:2 * level + 2]):
                    conv, bn = blk.layers[0], blk.layers[1]
                    s, t = _bn_affine(bn)
                    first = level == 0 and j == 0
                    w = pack_conv_weight(conv.weight.detach(), s, _CONV0_CHANNELS if first else None, 16 if first else None)
                    pair.append((w, t.float().contiguous()))
                self.layers.append(pair)


def _set_of(op):
    level, kind = divmod(op, 3)
    return 2 * level + 2 if kind == 2 else 2 * level + 1


class SparseTowers:
    """Both towers of a batch on the device.  Static buffers (capacity rows per voxel set), no host sync:
        levels_inp, levels_tmp = towers.run(points_inp, rgb_inp, points_tmp, rgb_tmp)
    each `levels_*` a list of four namespaces (.features (cap,C_l) fp32, .indices (cap,4) int32 bxyz; rows past the
    batch's count have batch id == B) — what Ops_GetPointFeat_spconv / Network.forward_from_backbone take."""

    def __init__(self, backbone_inp, backbone_tmp, device, batch, n_per, caps, unit=0.006, grid=64):
        if backbone_inp.training or backbone_tmp.training:
            raise ValueError("SparseTowers is an inference engine: put the towers in eval() mode (BatchNorm is folded)")
        self.device, self.b, self.n_per, self.unit, self.grid = device, batch, n_per, float(unit), grid
        self.caps = [int((c + 127) // 128 * 128) for c in caps]    # per set (9); set 0 unused
        assert len(self.caps) == NSETS
        lib = L.load()
        self.rpi = lib.dcl_spb_rows_per_instance()
        i32 = dict(dtype=torch.int32, device=device)
        self._src = (backbone_inp, backbone_tmp)
        self.repack()
        self.t = []
        for _ in range(2):
            t = types.SimpleNamespace()
            t.rows = torch.zeros(batch * self.rpi, dtype=torch.int64, device=device)
            t.prefix = torch.zeros(batch * self.rpi, **i32)
            t.counts = torch.zeros(NSETS * batch, **i32)
            t.offsets = torch.zeros(NSETS * (batch + 1), **i32)
            t.errors = torch.zeros(2, **i32)
            t.feat16 = torch.zeros(batch * n_per, 16, dtype=torch.float16, device=device)
            t.indices = [None] + [torch.zeros(self.caps[s], 4, **i32) for s in range(1, NSETS)]
            t.nbr = [torch.zeros(self.caps[_set_of(op)], 32, **i32) for op in range(NOPS)]
            t.anymask = [torch.zeros(self.caps[_set_of(op)] // 128, **i32) for op in range(NOPS)]
            t.conv16, t.subm32, t.pool32, t.pool16 = [], [], [], []
            for level in range(4):
                c_mid, c_out = DIMS[2 * level + 1], DIMS[2 * level + 2]
                t.conv16.append(torch.zeros(self.caps[2 * level + 1], c_mid, dtype=torch.float16, device=device))
                t.subm32.append(torch.zeros(self.caps[2 * level + 1], c_out, dtype=torch.float32, device=device))
                t.pool32.append(torch.zeros(self.caps[2 * level + 2], c_out, dtype=torch.float32, device=device))
                t.pool16.append(torch.zeros(self.caps[2 * level + 2], c_out, dtype=torch.float16, device=device)
                                if level < 3 else None)
            self.t.append(t)

    def repack(self):
        """(Re)build the packed weights from the tower modules (after load_state_dict / .to())."""
        self.weights = [_TowerWeights(bb) for bb in self._src]

    # ------------------------------------------------------------------ capacity planning (set-up time, syncs)
    @staticmethod
    def measure_counts(points_list, batch, n_per, device, unit=0.006, grid=64):
        """Rows per voxel set (9) of each given (B*n_per,3) cloud: run the set builder once and read the counts."""
        lib = L.load()
        rpi = lib.dcl_spb_rows_per_instance()
        out = []
        for pts in points_list:
            pts = pts.to(device).contiguous()
            rows = torch.zeros(batch * rpi, dtype=torch.int64, device=device)
            prefix = torch.zeros(batch * rpi, dtype=torch.int32, device=device)
            counts = torch.zeros(NSETS * batch, dtype=torch.int32, device=device)
            tin = (L.SpbTowerIn * 1)()
            tin[0].points, tin[0].rows, tin[0].prefix, tin[0].counts = L.ptr(pts), L.ptr(rows), L.ptr(prefix), L.ptr(counts)
            L.check(lib.dcl_spb_build_sets(batch, n_per, 1, ctypes.cast(tin, ctypes.c_void_p), unit, grid, L.stream_ptr()),
                    "spb_build_sets")
            out.append(counts.view(NSETS, batch).sum(1).tolist())
        return out

    @staticmethod
    def plan_capacities(points_list, batch, n_per, device, margin=1.3):
        counts = SparseTowers.measure_counts(points_list, batch, n_per, device)
        return [int(math.ceil(margin * max(c[s] for c in counts) / 128.0)) * 128 + 128 for s in range(NSETS)]

    # ------------------------------------------------------------------ one pass
    def run(self, points_inp, rgb_inp, points_tmp, rgb_tmp):
        lib, b, st = L.load(), self.b, L.stream_ptr
        ins = ((points_inp, rgb_inp), (points_tmp, rgb_tmp))
        tin = (L.SpbTowerIn * 2)()
        sets = (L.SpbTowerSets * 2)()
        for k, (t, (pts, rgb)) in enumerate(zip(self.t, ins)):
            assert pts.is_contiguous() and rgb.is_contiguous() and pts.shape == (b * self.n_per, 3) and rgb.shape == pts.shape
            tin[k].points, tin[k].rgb, tin[k].coords = L.ptr(pts), L.ptr(rgb), None
            tin[k].rows, tin[k].prefix, tin[k].counts = L.ptr(t.rows), L.ptr(t.prefix), L.ptr(t.counts)
            tin[k].feat16, tin[k].errors = L.ptr(t.feat16), L.ptr(t.errors)
            sets[k].rows, sets[k].prefix, sets[k].counts = L.ptr(t.rows), L.ptr(t.prefix), L.ptr(t.counts)
            sets[k].offsets, sets[k].errors = L.ptr(t.offsets), L.ptr(t.errors)
            for s in range(NSETS):
                sets[k].indices[s] = L.ptr(t.indices[s]).value if t.indices[s] is not None else None
                sets[k].cap[s] = self.caps[s]
            for op in range(NOPS):
                sets[k].nbr[op] = L.ptr(t.nbr[op]).value
                sets[k].anymask[op] = L.ptr(t.anymask[op]).value
        L.check(lib.dcl_spb_build_sets(b, self.n_per, 2, ctypes.cast(tin, ctypes.c_void_p), self.unit, self.grid, st()),
                "spb_build_sets")
        L.check(lib.dcl_spb_emit(b, 2, ctypes.cast(sets, ctypes.c_void_p), self.n_per, st()), "spb_emit")
        for level in range(4):
            c_in, c_mid, c_out = DIMS[2 * level], DIMS[2 * level + 1], DIMS[2 * level + 2]
            s_out = 2 * level + 1
            convs = (L.SpbConv * 2)()
            for k, t in enumerate(self.t):
                w, shift = self.weights[k].layers[level][0]
                src = t.feat16 if level == 0 else t.pool16[level - 1]
                self._fill_conv(convs[k], t, src, 3 * level, s_out, w, shift, t.conv16[level], None)
            ev = self._conv_event()
            L.check(lib.dcl_spb_conv3(b, 16 if level == 0 else c_in, c_mid, 2, ctypes.cast(convs, ctypes.c_void_p), st()),
                    "spb_conv3")
            self._conv_event(ev, 3 * level, c_in, c_mid)
            for k, t in enumerate(self.t):
                w, shift = self.weights[k].layers[level][1]
                self._fill_conv(convs[k], t, t.conv16[level], 3 * level + 1, s_out, w, shift, None, t.subm32[level])
            ev = self._conv_event()
            L.check(lib.dcl_spb_conv3(b, c_mid, c_out, 2, ctypes.cast(convs, ctypes.c_void_p), st()), "spb_conv3 (subm)")
            self._conv_event(ev, 3 * level + 1, c_mid, c_out)
            pools = (L.SpbPool * 2)()
            for k, t in enumerate(self.t):
                pools[k].inp, pools[k].nbr = L.ptr(t.subm32[level]), L.ptr(t.nbr[3 * level + 2])
                pools[k].offsets_out = t.offsets.data_ptr() + 4 * (2 * level + 2) * (b + 1)
                pools[k].out32, pools[k].out16 = L.ptr(t.pool32[level]), L.ptr(t.pool16[level])
                pools[k].cap_out = self.caps[2 * level + 2]
            L.check(lib.dcl_spb_avgpool(b, c_out, 2, ctypes.cast(pools, ctypes.c_void_p), st()), "spb_avgpool")
        return tuple([types.SimpleNamespace(features=t.pool32[level], indices=t.indices[2 * level + 2])
                      for level in range(4)] for t in self.t)

    @staticmethod
    def _conv_event(start=None, op=None, c_in=None, c_out=None):
        """CONV_EVENTS bookkeeping: called without arguments before a launch (returns the start event) and with it after."""
        if CONV_EVENTS is None:
            return None
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        if start is not None:
            CONV_EVENTS.append((start, ev, op, c_in, c_out))
        return ev

    def rulebook_pairs(self, op):
        """(output row, kernel offset) pairs with an input voxel in rulebook `op`, both towers: the multiply-accumulate
        count of that layer is pairs * c_in * c_out (slot 27 of the tile-transposed table holds each row's count)."""
        return int(sum(t.nbr[op].view(-1, 32, 128)[:, 27, :].sum().item() for t in self.t))

    def _fill_conv(self, slot, t, src16, op, s_out, w, shift, out16, out32):
        slot.in16, slot.nbr, slot.anymask = L.ptr(src16), L.ptr(t.nbr[op]), L.ptr(t.anymask[op])
        slot.offsets_out = t.offsets.data_ptr() + 4 * s_out * (self.b + 1)
        slot.w, slot.shift = L.ptr(w), L.ptr(shift)
        slot.out16, slot.out32, slot.cap_out = L.ptr(out16), L.ptr(out32), self.caps[s_out]

    def check_errors(self):
        """Host sync: raises if a point fell outside the voxel grid or a voxel set outgrew its buffer."""
        for name, t in zip(("inp", "tmp"), self.t):
            bad, over = t.errors.tolist()
            if over:
                raise RuntimeError(f"SparseTowers[{name}]: voxel sets {[s for s in range(NSETS) if over >> s & 1]} exceed "
                                   f"their capacities {self.caps}; rebuild with larger `caps`")
            if bad:
                raise RuntimeError(f"SparseTowers[{name}]: {bad} points outside the {self.grid}^3 voxel grid")

    def set_rows(self, tower, s):
        """(host sync, tests) number of rows of set s."""
        return int(self.t[tower].offsets.view(NSETS, self.b + 1)[s, self.b])


# ---------------------------------------------------------------------------------------- pointgroup_ops drop-ins
def voxelization_idx(coords, batch_size, mode=4):
    """libs/pointgroup_ops voxelization_idx (voxelize.cpp:10-34) on the device: coords (N,4) int bxyz with equal
    numbers of points per instance, grouped by instance -> (output_coords (M,4) int64, input_map (N,) int32,
    output_map (M, maxActive+1) int32) with the reference's first-appearance numbering.  The output shapes depend on
    the data, so this entry point synchronises (the towers themselves never do)."""
    assert mode == 4 and coords.is_cuda
    n = coords.shape[0]
    n_per = n // batch_size
    assert n_per * batch_size == n
    dev = coords.device
    lib = L.load()
    rpi = lib.dcl_spb_rows_per_instance()
    c32 = coords.to(torch.int32).contiguous()
    i32 = dict(dtype=torch.int32, device=dev)
    rows, prefix = torch.zeros(batch_size * rpi, dtype=torch.int64, device=dev), torch.zeros(batch_size * rpi, **i32)
    counts = torch.zeros(NSETS * batch_size, **i32)
    occupied, p2v = torch.zeros(n, 4, **i32), torch.zeros(n, **i32)
    v2p_sorted, v2p_start = torch.zeros(n, **i32), torch.zeros(n, **i32)
    tin = (L.SpbTowerIn * 1)()
    tin[0].coords, tin[0].rows, tin[0].prefix, tin[0].counts = L.ptr(c32), L.ptr(rows), L.ptr(prefix), L.ptr(counts)
    tin[0].occupied, tin[0].p2v, tin[0].v2p_sorted, tin[0].v2p_start = (L.ptr(x) for x in (occupied, p2v, v2p_sorted, v2p_start))
    L.check(lib.dcl_spb_build_sets(batch_size, n_per, 1, ctypes.cast(tin, ctypes.c_void_p), 0.006, 64, L.stream_ptr()),
            "voxelization_idx")
    cnt = counts[:batch_size].long()                                  # voxels per instance (host sync below)
    off = torch.cumsum(cnt, 0) - cnt
    keep = (torch.arange(n_per, device=dev)[None, :] < cnt[:, None]).reshape(-1)     # live slots
    output_coords = occupied[keep].long()
    inst = torch.arange(batch_size, device=dev).repeat_interleave(n_per)
    input_map = (p2v.long() + off[inst]).int()
    start = v2p_start.view(batch_size, n_per).long()
    end = torch.cat([start[:, 1:], torch.full((batch_size, 1), n_per, device=dev)], 1)
    end = torch.where(torch.arange(n_per, device=dev)[None, :] == (cnt[:, None] - 1), torch.full_like(end, n_per), end)
    num = (end - start).reshape(-1)[keep]
    max_active = int(num.max().item())
    m = output_coords.shape[0]
    output_map = torch.zeros(m, max_active + 1, **i32)
    output_map[:, 0] = num.int()
    base = (start + (torch.arange(batch_size, device=dev) * n_per)[:, None]).reshape(-1)[keep]      # into v2p_sorted
    glob = v2p_sorted.long() + inst * n_per                                                          # global point index
    for j in range(max_active):
        live = num > j
        output_map[live, 1 + j] = glob[base[live] + j].int()
    return output_coords, input_map, output_map


def voxelization(feats, map_rule, mode=4):
    """libs/pointgroup_ops voxelization (voxelize.cu:10-31), mode 4: feats (N,C) fp32, map_rule (M, maxActive+1) int32
    -> (M,C) means, summed in rule order with the 1/n multiplier applied first (bit-exact with the reference)."""
    assert mode == 4
    feats = L.require(feats.contiguous(), torch.float32, "feats")
    rules = L.require(map_rule.contiguous(), torch.int32, "map_rule")
    m, width = rules.shape
    out = torch.empty(m, feats.shape[1], dtype=torch.float32, device=feats.device)
    L.check(L.load().dcl_voxelize_mean(m, width, feats.shape[1], L.ptr(feats), L.ptr(rules), L.ptr(out), L.stream_ptr()),
            "voxelization")
    return out
