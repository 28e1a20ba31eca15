"""Layer-0 activations of the first decode step (decode_debug dump) against the oracle's modules."""
import sys, os, ctypes as C
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import synthetic
from tests.gpu_util import make_b200_model
torch.set_printoptions(linewidth=200, precision=3, sci_mode=False)
o = synthetic.make_model(stop_bias=-8.0)
g = make_b200_model(o)
DR = int(os.environ.get("DUMP_RANK", "0")); g.set_option("decode_debug", 1 + DR)
B, S, T = int(sys.argv[1]) if len(sys.argv) > 1 else 1, int(sys.argv[2]) if len(sys.argv) > 2 else 10, 1
ph, pl, _, _ = synthetic.make_inputs(B, S, 8, 41, ragged=True)
g.inference(ph.cuda(), pl.cuda(), max_len=T, seed=7)
def dump(slot, n):
    out = torch.zeros(B * n)
    g._check(g._lib.tts_debug_read_dump(g._handle, g._ws.data_ptr(), slot * 2560, B * n, C.c_void_p(out.data_ptr()), g._stream()), "dump")
    return out.view(B, n)
with torch.no_grad():
    b_ids = np.arange(B)
    mem = o.encode(ph, pl)
    st = o.init_decode_state(mem, pl, T)
    frame = torch.zeros(B, 1, 80)
    x = o._dec_prenet(frame, 7, np.array([0]), b_ids) + o.dec_alpha * o.pe[0][None, None]
    def cmp(name, ref, got):
        ref = ref.reshape(B, -1); print(f"{name:14s} rel err {float((got - ref).norm() / ref.norm()):.4f}   max abs {float((got - ref).abs().max()):.4f}")
    cmp("prenet x", x, dump(0, 512))
    layer = o.decoder.layers[0]; sa = layer.self_attn
    q, k, v = sa.wq(x), sa.wk(x), sa.wv(x)
    d1 = dump(1, 192)           # head 0 of rank 0: q | k | v (64 each)
    cmp("q head0", q[:, 0, :64], d1[:, :64]); cmp("k head0", k[:, 0, :64], d1[:, 64:128]); cmp("v head0", v[:, 0, :64], d1[:, 128:])
    a_pre = v[:, 0]             # t = 0: softmax over one key -> attention output before wo is v itself
    cmp("self ctx", a_pre, dump(2, 512))
    a = sa.wo(a_pre[:, None])
    x1 = layer.norm1(x + a); cmp("LN1", x1, dump(3, 512))
    ca = layer.cross_attn
    ctx = ca.attend(ca.split(ca.wq(x1)), st["ck"][0], st["cv"][0], st["cross_mask"])     # includes wo?
    import inspect; print(inspect.getsource(ca.attend)[:600])
    cmp("LN2 (via oracle)", layer.norm2(x1 + ctx), dump(5, 512))
    x2 = layer.norm2(x1 + ctx)
    h = torch.relu(layer.ffn.w1(x2)); cmp("ffn hidden r0", h[:, 0, :256], dump(6, 256))
    x3 = layer.norm3(x2 + layer.ffn(x2)); cmp("LN3", x3, dump(7, 512))
    w2 = layer.ffn.w2
    part0 = h[:, 0, :256] @ w2.weight[:, :256].T
    cmp("FFN2 partial r0", part0, dump(8, 512))
    cmp("pre-LN3 y", x2 + layer.ffn(x2), dump(9, 512))
    d = (dump(9, 512) - (x2 + layer.ffn(x2)).reshape(B, -1)).abs()[0]
    print("pre-LN3 abs err per 64-col slice:", d.view(8, 64).mean(1))
    d = (dump(8, 512) - part0.reshape(B, -1)).abs()[0]
    print("partial abs err per 32-col (warp) slice:", d.view(16, 32).mean(1))
    if DR:
        hfull = torch.relu(layer.ffn.w1(x2))[:, 0]            # [B, 2048]
        rv = dump(10, 2560).view(8, 5, 64)[:, 0]              # [sender rank][64] for utterance 0 (B = 1)
        for r in range(8):
            ref = hfull[0, 256 * r:256 * r + 256] @ w2.weight[64 * DR:64 * DR + 64, 256 * r:256 * r + 256].T
            print("recv from rank", r, "max abs err", float((rv[r] - ref).abs().max()), " |ref|max", float(ref.abs().max()))
