import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gemini_b200 as gm
from gemini_b200 import snark
from gemini_b200.devvec import DeviceFr
from gemini_b200.transcript import MerlinTranscript
R = gm.field.R
ctx = gm.Context(0)
logn = 24
n = 1 << logn
srs = ctx.srs_generate(n, first_multiple=1)
srs.precompute()
ctx.synchronize()
ck = gm.CommitterKey(ctx, srs)
out = {"msm_sweep_ms": {}}
for k in list(range(1, 25)):
    v = DeviceFr.random(ctx, 1 << k, 100 + k)
    best = 1e9
    for rep in range(3):
        ctx.synchronize(); t0 = time.perf_counter()
        ck.commit(v)
        ctx.synchronize(); best = min(best, time.perf_counter() - t0)
    out["msm_sweep_ms"][k] = round(best * 1e3, 3)
    v.free()
print(json.dumps(out))
# tensorcheck pieces on the prover's own vectors
e = 0x1234567890ABCDEF1234567890ABCDEF1234567890ABCDEF % R
r1cs = snark.R1cs.dummy(ctx, n, e)
abc = DeviceFr.random(ctx, n, 7)
chals = [(i * 7919 + 13) % R for i in range(logn)]
def T(name, f, acc):
    ctx.synchronize(); t0 = time.perf_counter(); r = f(); ctx.synchronize(); acc[name] = round((time.perf_counter() - t0) * 1e3, 3); return r
for rep in range(2):
    acc = {}
    def mk_lc():
        lc = DeviceFr.zeros(ctx, n); lc.axpy(1, abc); lc.axpy(5, r1cs.z); return lc
    lc = T("lincomb", mk_lc, acc)
    foldings = T("fold_chain", lambda: lc.fold_chain(chals[:-1]), acc)
    comms = T("batch_commit_23", lambda: ck.batch_commit(foldings), acc)
    T("serial_commit_23", lambda: [ck.commit(f) for f in foldings], acc)
    T("evals_base", lambda: [(r1cs.w.evaluate(12345), r1cs.w.evaluate_pm(777))], acc)
    T("evals_foldings", lambda: [f.evaluate_pm(777) for f in foldings], acc)
    def opening():
        polys = [r1cs.w] + foldings
        b = DeviceFr.zeros(ctx, n); eta = 1
        for p in polys:
            b.axpy(eta, p); eta = eta * 31337 % R
        return b
    b = T("eta_combination", opening, acc)
    def divs():
        q = b
        for pt in (5, 7, R - 7):
            q, _ = q.div_linear(pt)
        return q
    q = T("three_divisions", divs, acc)
    T("commit_quotient", lambda: ck.commit(q), acc)
    T("commit_w_allequal", lambda: ck.commit(r1cs.w), acc)
    print(json.dumps({"tensorcheck_pieces_ms": acc}))
timers = {}
snark.new_time(ctx, r1cs, ck, MerlinTranscript(), timers)
timers = {}
ctx.synchronize(); t0 = time.perf_counter()
snark.new_time(ctx, r1cs, ck, MerlinTranscript(), timers)
ctx.synchronize()
print(json.dumps({"new_time_s": time.perf_counter() - t0, "phases": timers}))
