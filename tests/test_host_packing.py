"""Host-side packing of the tensor-core path, checked on the CPU: the packed weight image follows the byte formula
of include/dcl_b200.h, eval-mode BatchNorm folding reproduces the layer modules, and the refiner's first layer is
permuted / padded the way FusedRefiner feeds it."""
import torch

from dcl_net_b200 import fused_tail as FT
from dcl_net_b200.modules import BasicBlock_3DCONV, Head_MultiLayerPerceptron
from dcl_net_b200.refiner import Refiner


def unpack_weight(packed, cout, cin, nt):
    """Inverse of pack_weight through the documented map:
    byte(o,i,half) = ((o/nt)(cin/32) + i/32)(nt*128) + half*(nt*64) + ((o%nt)/8)*512 + ((i%32)/8)*128 + (o%8)*16 + (i%8)*2"""
    raw = packed.view(torch.int16)          # bf16 bit patterns, 2 bytes each
    o = torch.arange(cout).view(-1, 1).expand(cout, cin)
    i = torch.arange(cin).view(1, -1).expand(cout, cin)
    base = ((o // nt) * (cin // 32) + i // 32) * (nt * 128) + ((o % nt) // 8) * 512 + ((i % 32) // 8) * 128 + (o % 8) * 16 + (i % 8) * 2
    out = torch.zeros(cout, cin, dtype=torch.float64)
    for half in (0, 1):
        bits = raw[((base + half * nt * 64) // 2).reshape(-1)].reshape(cout, cin)
        out += bits.view(torch.bfloat16).double()
    return out


def test_pack_weight_follows_the_header_formula():
    g = torch.Generator().manual_seed(0)
    for cout, cin in ((256, 480), (128, 256), (64, 32), (1024, 512)):
        w = torch.randn(cout, cin, generator=g)
        nt = FT.pick_nt(cout)
        packed = FT.pack_weight(w, nt)
        assert packed.numel() == cout * cin * 4
        back = unpack_weight(packed, cout, cin, nt)
        assert (back - w.double()).abs().max().item() <= 2.0 ** -16 * w.abs().max().item()   # hi + lo carries 16 bits


def _randomise_bn(mod, g):
    for m in mod.modules():
        if isinstance(m, (torch.nn.BatchNorm1d, torch.nn.BatchNorm3d)):
            m.running_mean.copy_(0.2 * torch.randn(m.num_features, generator=g))
            m.running_var.copy_(0.5 + torch.rand(m.num_features, generator=g))
            m.weight.data.copy_(0.5 + torch.rand(m.num_features, generator=g))
            m.bias.data.copy_(0.2 * torch.randn(m.num_features, generator=g))


def _apply(layer, x):
    w = unpack_weight(layer.w, layer.cout, layer.cin, layer.nt)
    y = x.double() @ w.T + (layer.bias.double() if layer.bias is not None else 0.0)
    if layer.relu:
        y = y.clamp_min(0)
    if layer.post_scale is not None:
        y = y * layer.post_scale.double() + layer.post_shift.double()
    return y


def test_disengage_stack_folding_matches_the_modules():
    g = torch.Generator().manual_seed(1)
    torch.manual_seed(1)
    stack = torch.nn.Sequential(BasicBlock_3DCONV(480, 256, False, 1, 1, 0, True, "relu", 0.0),
                                BasicBlock_3DCONV(256, 128, False, 1, 1, 0, True, "relu", 0.0)).eval()
    _randomise_bn(stack, g)
    x = torch.randn(40, 480, generator=g)
    with torch.no_grad():
        want = stack(x.T.reshape(1, 480, 40, 1, 1)).reshape(128, 40).T.double()
        y = x
        for layer in FT.layers_from_disengage(stack):
            y = _apply(layer, y).float()
    assert (y.double() - want).abs().max().item() <= 1e-4 * want.abs().max().item()


def test_head_folding_relu_then_batchnorm():
    g = torch.Generator().manual_seed(2)
    torch.manual_seed(2)
    head = Head_MultiLayerPerceptron([512, 512, 512, 1024], ["relu"] * 3, [True] * 3, [0.0] * 3).eval()
    _randomise_bn(head, g)
    x = torch.randn(24, 512, generator=g)
    with torch.no_grad():
        want = head(x.T.unsqueeze(0)).squeeze(0).T.double()
        gemm, rest = FT.layers_from_head(head)
        assert len(gemm) == 3 and not rest
        y = x
        for layer in gemm:
            y = _apply(layer, y).float()
    assert (y.double() - want).abs().max().item() <= 1e-4 * want.abs().max().item()


def test_refiner_first_layer_permutation():
    torch.manual_seed(3)
    ref = Refiner().eval()
    fused = FT.FusedRefiner(ref)
    l1 = fused.layers[0]
    assert (l1.cin, l1.cout) == (288, 512)
    w = unpack_weight(l1.w, 512, 288, l1.nt)
    conv = [m for m in ref.MLP_share.layers if isinstance(m, torch.nn.Conv1d)][0]
    w_ref = conv.weight.detach().reshape(512, 259).double()
    tol = 2.0 ** -16 * w_ref.abs().max().item()
    assert (w[:, :256] - w_ref[:, 3:]).abs().max().item() <= tol       # F_Xo_p channels first
    assert (w[:, 256:259] - w_ref[:, :3]).abs().max().item() <= tol    # then x, y, z
    assert w[:, 259:].abs().max().item() == 0.0                        # zero padding to a whole k-block


def test_split_product_precision_rule():
    """The precision rule of every tensor-core contraction here: x = hi + lo in bf16, products
    hi*hi + hi*lo + lo*hi accumulated in fp32.  Emulated on the CPU: the result is fp32-faithful (normwise error
    ~1e-5 or better) where plain bf16 operands miss the 1e-3 bar by an order of magnitude on unscaled post-ReLU dot
    products."""
    g = torch.Generator().manual_seed(7)
    x = torch.randn(256, 480, generator=g).relu()
    w = torch.randn(256, 480, generator=g) / 480 ** 0.5

    def split(t):
        hi = t.to(torch.bfloat16).float()
        return hi, (t - hi).to(torch.bfloat16).float()
    xh, xl = split(x)
    wh, wl = split(w)
    exact = x.double() @ w.double().T
    three = (xh @ wh.T + xh @ wl.T + xl @ wh.T).double()          # fp32 accumulation, as the TMEM accumulator
    plain = (xh @ wh.T).double()
    scale = exact.abs().max().item()
    assert (three - exact).abs().max().item() <= 2e-5 * scale
    assert (plain - exact).abs().max().item() >= 1e-3 * scale     # why a single bf16 product is not enough
    # logits of the FDA: unscaled dot products of post-ReLU features, |S| ~ 40 -> softmax weights
    q, k = torch.randn(128, 128, generator=g).relu(), torch.randn(1024, 128, generator=g).relu()
    qh, ql = split(q)
    kh, kl = split(k)
    a_exact = torch.softmax(q.double() @ k.double().T, dim=1)
    a_three = torch.softmax((qh @ kh.T + qh @ kl.T + ql @ kh.T).double(), dim=1)
    a_plain = torch.softmax((qh @ kh.T).double(), dim=1)
    rel = lambda a: ((a - a_exact).abs().max() / a_exact.abs().max()).item()
    assert rel(a_three) < 1e-3 and rel(a_plain) > 1e-2


def _split(t):
    hi = t.to(torch.bfloat16).float()
    return hi, (t - hi).to(torch.bfloat16).float()


def _mm3(a, b):
    """hi*hi + hi*lo + lo*hi with fp32 accumulation (what three tcgen05.mma into one TMEM accumulator compute)."""
    ah, al = _split(a)
    bh, bl = _split(b)
    return ah @ bh + ah @ bl + al @ bh


def _fda_emulated(ri1, ri2, re2, key_block=64, threshold=8.0):
    """CPU emulation of the fused FDA kernel's arithmetic (csrc/fda.cu) for one instance: 64-key blocks, online
    softmax in log2 units with the LAZY rescale (the running reference maximum only moves when a block exceeds it by
    more than `threshold`), P rounded to bf16 hi + lo before the value product, the row sum taken over the rounded
    weights, fp32 accumulators.  ri1 (C,N) queries, ri2 (C,M) keys, re2 (P,M) values -> (P,N), (C,N)."""
    log2e = 1.4426950408889634
    n, m = ri1.shape[1], ri2.shape[1]
    v = torch.cat([re2, ri2], 0)                                   # value rows: RE_2 then RI_2
    o = torch.zeros(n, v.shape[0])
    l = torch.zeros(n)
    m_ref = torch.full((n,), float("-inf"))
    for j0 in range(0, m, key_block):
        s = _mm3(ri1.T.contiguous(), ri2[:, j0:j0 + key_block].contiguous())            # (n, 64) fp32
        mx = s.max(dim=1).values * log2e
        if j0 == 0:
            m_ref = mx.clone()
            alpha = torch.ones(n)
        else:
            need = mx > m_ref + threshold
            alpha = torch.where(need, torch.exp2(m_ref - mx), torch.ones(n))
            m_ref = torch.where(need, mx, m_ref)
        p = torch.exp2(s * log2e - m_ref[:, None])
        ph, pl = _split(p)
        l = l * alpha + (ph + pl).sum(dim=1)
        o = o * alpha[:, None] + _mm3(p, v[:, j0:j0 + key_block].T.contiguous())
    out = (o / l[:, None]).T
    return out[:re2.shape[0]], out[re2.shape[0]:]


def test_fda_arithmetic_emulated_on_cpu():
    """The numerical design of the fused FDA kernel meets the 1e-3 bar with margin on the CPU emulation: network-like
    inputs, near-one-hot softmax (|logit| ~ 300), and key norms that keep growing (the lazy rescale path)."""
    g = torch.Generator().manual_seed(11)
    c, n, m, p = 64, 128, 512, 256
    cases = {
        "relu": (torch.randn(c, n, generator=g).relu(), torch.randn(c, m, generator=g).relu(), 1e-4),
        "peaked": (3.0 * torch.randn(c, n, generator=g), 3.0 * torch.randn(c, m, generator=g), 5e-4),
        "growing": (torch.randn(c, n, generator=g).relu(),
                    torch.randn(c, m, generator=g).relu() * torch.linspace(0.2, 2.5, m)[None, :], 1e-4),
    }
    for name, (ri1, ri2, tol) in cases.items():
        re2 = torch.randn(p, m, generator=g)
        a = torch.softmax(ri2.double().T @ ri1.double(), dim=0)                          # (m, n), softmax over keys
        want_e, want_i = re2.double() @ a, ri2.double() @ a
        got_e, got_i = _fda_emulated(ri1, ri2, re2)
        for got, want in ((got_e, want_e), (got_i, want_i)):
            err = (got.double() - want).abs().max().item() / want.abs().max().item()
            assert err <= tol, (name, err)


def test_pm_image_equals_packed_weights_of_ntile_128():
    """What the training path's weight-gradient GEMM rests on (train_tail.py, csrc/train_ops.cu): the PM image of an
    (R x C) matrix — byte(r,c,half) = ((r/128)(C/32) + c/32)*16384 + half*8192 + ((r%128)/8)*512 + ((c%32)/8)*128 +
    (r%8)*16 + (c%8)*2 — is byte for byte the packed-weight layout with n-tile 128, so an operand image written by
    the tile pass can be read by dcl_pm_gemm as its `w` operand."""
    g = torch.Generator().manual_seed(1)
    rows, c = 384, 96
    x = torch.randn(rows, c, generator=g)
    packed = FT.pack_weight(x, 128).view(torch.int16)
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)
    r = torch.arange(rows).view(-1, 1).expand(rows, c)
    ch = torch.arange(c).view(1, -1).expand(rows, c)
    base = ((r // 128) * (c // 32) + ch // 32) * 16384 + ((r % 128) // 8) * 512 + ((ch % 32) // 8) * 128 + (r % 8) * 16 + (ch % 8) * 2
    for half, img in ((0, hi), (1, lo)):
        got = packed[((base + half * 8192) // 2).reshape(-1)].reshape(rows, c)
        assert torch.equal(got, img.view(torch.int16))
