"""Development check of the tensor-core block pass (tc_kernels.cu): parity against the oracle and against
the all-fp64 path on small shapes, then timings on a C4-shaped shard for several block sizes."""
import os, sys, time
import numpy as np, scipy.sparse as sp
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import plaid_b200 as pb
from plaid_b200 import api, synth
from oracle import plaid_oracle as O

def relmax(a, b):
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))

def check(tag, X, G, names, fn_gpu, fn_ora, **kw):
    ctx = pb.Context(0)
    api.EXACT_FP64 = False
    got = fn_gpu(pb.NamedMatrix(X, names), pb.NamedMatrix(G, names), ctx=ctx, **kw).mat
    info = ctx.plan_info()
    api.EXACT_FP64 = True
    ex = fn_gpu(pb.NamedMatrix(X, names), pb.NamedMatrix(G, names), ctx=ctx, **kw).mat
    api.EXACT_FP64 = False
    want = fn_ora(O.Named(X, names), O.Named(G, names), **kw).mat
    print(f"{tag}: tc vs oracle {relmax(got, want):.3e} | fp64 vs oracle {relmax(ex, want):.3e} | plan {info}", flush=True)
    ctx.close()

if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "parity"
    if what == "parity":
        # dense: every row goes through the tensor cores (P padded 1000 -> 1024, S 500 -> 4 tiles, N ragged)
        P, N, S = 1000, 203, 500
        Xd = synth.dense_x_numpy(P, N, seed=5)
        G = synth.genesets_numpy(P, S, seed=6, size_cap=(5, 300))
        names = synth.gene_names(P)
        check("dense plaid raw", Xd, G, names, pb.plaid, O.plaid, normalize=False)
        check("dense plaid norm", Xd, G, names, pb.plaid, O.plaid)
        check("dense gsva", Xd, G, names, pb.replaid_gsva, O.replaid_gsva)
        check("dense ssgsea", Xd, G, names, pb.replaid_ssgsea, O.replaid_ssgsea)
        # sparse: block of high-degree rows + scatter
        P, N, S = 3000, 517, 2500
        X = synth.sparse_x_numpy(P, N, seed=7)
        G = synth.genesets_numpy(P, S, seed=8, size_cap=(5, 400))
        names = synth.gene_names(P)
        check("sparse plaid raw", X, G, names, pb.plaid, O.plaid, normalize=False)
        check("sparse plaid norm", X, G, names, pb.plaid, O.plaid)
        check("sparse sing", X, G, names, pb.replaid_sing, O.replaid_sing)
        check("sparse ssgsea", X, G, names, pb.replaid_ssgsea, O.replaid_ssgsea)
        check("sparse ucell", X, G, names, pb.replaid_ucell, O.replaid_ucell)
        check("sparse scse", X, G, names, pb.replaid_scse, O.replaid_scse)
        # a NaN in a block row: the flag routes the block through the fp64 gather passes
        X2 = X.copy().tocsc(); X2.data[X2.indices == 0] = np.nan
        ctx = pb.Context(0)
        got = pb.plaid(pb.NamedMatrix(X2, names), pb.NamedMatrix(G, names), normalize=False, ctx=ctx).mat
        want = O.plaid(O.Named(X2, names), O.Named(G, names), normalize=False).mat
        ok = np.array_equal(np.isnan(got), np.isnan(want))
        m = ~np.isnan(want)
        print(f"sparse NaN fallback: nan pattern equal {ok}, finite err {relmax(got[m], want[m]):.3e}", flush=True)
    else:
        import torch, bench
        from plaid_b200 import _lib as L
        Nc = int(os.environ.get("CELLS", "32768"))
        G, xp, xi, xx = bench.make_inputs("cuda:0", 0, Nc)
        names = synth.gene_names(bench.P_GENES)
        out = torch.empty(bench.S_SETS * Nc, dtype=torch.float64, device="cuda:0")
        Xn = pb.NamedMatrix(pb.DeviceCSC(xp, xi, xx, (bench.P_GENES, Nc)), names)
        Gn = pb.NamedMatrix(G, names)
        ref = None
        for cfg in os.environ.get("CFGS", "off,1024,2048,3072,4096").split(","):
            if cfg == "auto":
                os.environ["PLAIDGPU_TC"] = "1"
                os.environ.pop("PLAIDGPU_TC_K", None)
            elif cfg == "off":
                os.environ["PLAIDGPU_TC"] = "0"
            else:
                os.environ["PLAIDGPU_TC"] = "1"
                os.environ["PLAIDGPU_TC_K"] = cfg
            ctx = pb.Context(0)
            for fn, nm in ((pb.plaid, "plaid"), (pb.replaid_ssgsea, "ssgsea")):
                for it in range(3):
                    fn(Xn, Gn, ctx=ctx, out=out)
                torch.cuda.synchronize()
                ms = [round(ctx.kernel_ms(k), 2) for k in range(4)]
                chk = float(out[: bench.S_SETS * 64].double().abs().sum().item())
                print(f"cfg {cfg} {nm} cells {Nc}: score {ms[0]} colstats {ms[1]} fixup {ms[2]} rank {ms[3]} ms | plan {ctx.plan_info()} | checksum {chk:.12e}", flush=True)
            ctx.close()
