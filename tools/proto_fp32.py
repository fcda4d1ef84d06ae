"""Developer tool (not a test, not product): numpy FP32 model of the device numerics of k_align3 -- per-DIAGONAL scalar
offsets (integer re-basing every K diagonals) and a fixed tilt g along x -- checked against the oracle, to decide the
scheme before spending GPU time.   python tools/proto_fp32.py --lx 1500 --e 64 --K 4 --g 0"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cpecan-signal_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracleshim as O  # noqa: E402
import parity  # noqa: E402
from cpecan_signal import synth  # noqa: E402

f32 = np.float32
NI = f32(-np.inf)
SEG = [(1.0, -0.009350833524763, 0.130659527668286, 0.498799810682272, 0.693203116424741),
       (2.5, -0.014532321752540, 0.139942324101744, 0.495635523139337, 0.692140569840976),
       (4.5, -0.004605031767994, 0.063427417320019, 0.695956496475118, 0.514272634594009),
       (7.5, -0.000458661602210, 0.009695946122598, 0.930734667215156, 0.168037164329057)]
TAB = np.zeros((17, 4), dtype=np.float32)
for i in range(16):
    a = i / 2.0
    for hi, c3, c2, c1, c0 in SEG:
        if a <= hi:
            TAB[i] = [f32(c3), f32(c2), f32(f32(c1) - f32(1.0)), f32(c0)]
            break


DT = np.float32
FWD64 = False


def la(x, y):
    with np.errstate(invalid="ignore"):
        m = np.maximum(x, y)
        a = np.abs(x - y)
        a = np.where(np.isnan(a), DT(8.0), np.minimum(a, DT(8.0))).astype(DT)
        i = np.ceil(a * DT(2.0)).astype(np.int64)
        c = TAB[i].astype(DT)
        u = a * c[:, 0] + c[:, 1]
        v = a * c[:, 2] + c[:, 3]
        q = (u * (a * a) + v).astype(DT)
        return (m + q).astype(DT)


def kmer_codes(ref, lX):
    b = np.frombuffer(ref.encode(), dtype=np.uint8)
    code = np.full(len(b), -1, dtype=np.int64)
    for ch, v in zip(b"ACGT", range(4)):
        code[b == ch] = v
    k = np.zeros(lX, dtype=np.int64)
    for j in range(6):
        k = k * 4 + code[j:j + lX]
    return k


def run(read, tables, e, K, g_, expanded, cwin=-1, ridgeB=0, precise=0, emis64=0, thr=0.01, minDiags=1000, tbDiags=40):
    global DT
    g = g_
    l1, l2, l3 = tables
    lX, lY = read.lX, read.lY
    D = lX + lY
    band = O.band(read.anchors, lX, lY, e)
    lo = ((band[:, 0] + band[:, 1]) // 2).astype(np.int64)
    hi = ((band[:, 0] + band[:, 2]) // 2).astype(np.int64)
    match = synth.scale_match_table(l1, read.scale5)[1:].reshape(4096, 5)
    gapy = l3[1:].reshape(4096, 5)
    k = kmer_codes(read.ref, lX)
    c0 = 68.0 * read.scale5[0] + read.scale5[1]
    HL2PI = 0.91893853320467267
    # column records (index = matrix x, 0 .. lX+1)
    W = lX + 2
    mu_m = np.zeros(W); c1m = np.zeros(W); nu_m = np.zeros(W); q_m = np.zeros(W); K_m = np.full(W, -np.inf)
    mu_y = np.zeros(W); c1y = np.zeros(W); nu_y = np.zeros(W); q_y = np.zeros(W); K_y = np.full(W, -np.inf)
    eXc = np.full(W, -np.inf)
    xs = np.arange(1, lX + 1)
    mm = match[k]; gy = gapy[k]
    mu_m[xs] = mm[:, 0] - c0; c1m[xs] = -0.5 / mm[:, 1] ** 2; nu_m[xs] = mm[:, 2]; q_m[xs] = -0.5 / mm[:, 3] ** 2
    K_m[xs] = -2 * HL2PI - np.log(mm[:, 1]) - np.log(mm[:, 3])
    mu_y[xs] = gy[:, 0] - c0; c1y[xs] = -0.5 / gy[:, 1] ** 2; nu_y[xs] = gy[:, 2]; q_y[xs] = -0.5 / gy[:, 3] ** 2
    K_y[xs] = -2 * HL2PI - np.log(gy[:, 1]) - np.log(gy[:, 3])
    eXc[xs] = np.log(0.1)
    cols = [a.astype(np.float32) for a in (mu_m, c1m, nu_m, q_m, K_m, mu_y, c1y, nu_y, q_y, K_y, eXc)]
    mu_m, c1m, nu_m, q_m, K_m, mu_y, c1y, nu_y, q_y, K_y, eXc = cols
    if expanded:
        A_m = (K_m.astype(np.float64) + c1m.astype(np.float64) * mu_m.astype(np.float64) ** 2 + q_m.astype(np.float64) * nu_m.astype(np.float64) ** 2).astype(np.float32)
        B_m = (-2.0 * c1m.astype(np.float64) * mu_m).astype(np.float32); D_m = (-2.0 * q_m.astype(np.float64) * nu_m).astype(np.float32)
        A_y = (K_y.astype(np.float64) + c1y.astype(np.float64) * mu_y.astype(np.float64) ** 2 + q_y.astype(np.float64) * nu_y.astype(np.float64) ** 2).astype(np.float32)
        B_y = (-2.0 * c1y.astype(np.float64) * mu_y).astype(np.float32); D_y = (-2.0 * q_y.astype(np.float64) * nu_y).astype(np.float32)
    evm = np.zeros(lY + 1, dtype=np.float32); evn = np.ones(lY + 1, dtype=np.float32)
    evm[1:] = (read.events[:, 0] - c0).astype(np.float32); evn[1:] = read.events[:, 1].astype(np.float32)
    evm2 = (evm * evm).astype(np.float32); evn2 = (evn * evn).astype(np.float32)

    def emissions(d, x):          # arrays over x (matrix columns), y = d - x
        y = d - x
        m, n = evm[y], evn[y]
        if expanded:
            eM = (A_m[x] + B_m[x] * m + c1m[x] * evm2[y] + D_m[x] * n + q_m[x] * evn2[y]).astype(np.float32)
            eY = (A_y[x] + B_y[x] * m + c1y[x] * evm2[y] + D_y[x] * n + q_y[x] * evn2[y]).astype(np.float32)
        elif emis64:
            dm = m.astype(np.float64) - mu_m[x]; dn = n.astype(np.float64) - nu_m[x]
            eM = (c1m[x].astype(np.float64) * (dm * dm) + (q_m[x].astype(np.float64) * (dn * dn) + K_m[x])).astype(np.float32)
            dm = m.astype(np.float64) - mu_y[x]; dn = n.astype(np.float64) - nu_y[x]
            eY = (c1y[x].astype(np.float64) * (dm * dm) + (q_y[x].astype(np.float64) * (dn * dn) + K_y[x])).astype(np.float32)
        else:
            dm = m - mu_m[x]; dn = n - nu_m[x]
            eM = (c1m[x] * (dm * dm) + (q_m[x] * (dn * dn) + K_m[x])).astype(np.float32)
            dm = m - mu_y[x]; dn = n - nu_y[x]
            eY = (c1y[x] * (dm * dm) + (q_y[x] * (dn * dn) + K_y[x])).astype(np.float32)
        return eM, eY, eXc[x]

    t = [f32(v) for v in O.NANOPORE_TRANSITIONS]
    tMC, tMX, tMY, tOX, tOY, tEX, tEY = t[0], t[1], t[2], t[3], t[4], t[5], t[6]
    g = f32(g)
    # ragged (1,1)
    startv = (NI, f32(0), f32(0))
    endv = (tMC, tMX, tMY)
    rendv = (f32((float(O.NANOPORE_TRANSITIONS[3]) + float(O.NANOPORE_TRANSITIONS[4])) / 2.0), tEX, tEY)

    # full-width diagonals: arrays over x = 0 .. lX (index x+1 so that x-1 = -1 is valid) -> index = x + 1
    def blank(dt=None):
        return [np.full(lX + 3, NI, dtype=dt or DT) for _ in range(3)]

    rows = {}       # d -> (lo, hi, M, X, Y, eM, eY, eX, U)
    stats = dict(maxabs_ridgeF=0.0, maxabs_ridgeB=0.0)
    pairs = []
    totals = np.full(D + 1, np.nan)

    def measure(vals, lo_, hi_, Frow=None):
        if cwin >= 0:
            if Frow is not None and ridgeB:
                j = int(np.argmax(Frow + vals[0][lo_ + 1:hi_ + 2]))
                m = float(vals[0][lo_ + 1 + j])
                if np.isfinite(m):
                    return int(np.rint(m))
            xc = (lo_ + hi_) // 2
            sl = slice(max(lo_, xc - cwin) + 1, min(hi_, xc + cwin) + 2)
            m = max(float(vals[0][sl].max()), float(vals[1][sl].max()), float(vals[2][sl].max()))
            if np.isfinite(m):
                return int(np.rint(m))
        sl = slice(lo_ + 1, hi_ + 2)
        m = max(float(vals[0][sl].max()), float(vals[1][sl].max()), float(vals[2][sl].max()))
        return 0 if not np.isfinite(m) else int(np.rint(m))

    # forward state
    F1 = blank(np.float64 if FWD64 else None); F2 = blank(np.float64 if FWD64 else None)      # d-1, d-2
    F1[0][1], F1[1][1], F1[2][1] = startv
    U1 = 0; U2 = 0                   # units of d-1, d-2
    rows[0] = (0, 0, F1[0][1:2].copy(), F1[1][1:2].copy(), F1[2][1:2].copy(), None, None, None, 0)
    tracedBackTo = 0
    dcur = 0
    pend = 0                         # pending integer shift for the next diagonal
    while tracedBackTo < D:
        # ---------------- forward
        d = dcur
        while True:
            d += 1
            l, h = lo[d], hi[d]
            U = U1 + pend; pend = 0
            x = np.arange(l, h + 1)
            eM, eY, eX = emissions(d, x)
            s1 = DT(U1 - U); s2 = DT(U2 - U)
            L = [a[x] for a in F1]          # index x  == (x-1)+1
            Mi = [a[x] for a in F2]
            own = [a[x + 1] for a in F1]
            if FWD64:
                DTsave = DT; DT = np.float64
            tX = la(L[0] + (tOX + s1 - g), L[1] + (tEX + s1 - g))
            tM = la(la(Mi[0] + (tMC + s2 - g), Mi[1] + (tMX + s2 - g)), Mi[2] + (tMY + s2 - g))
            tY = la(own[0] + (tOY + s1), own[2] + (tEY + s1))
            if FWD64:
                DT = DTsave
            new = blank(np.float64 if FWD64 else None)
            new[0][x + 1] = tM + eM; new[1][x + 1] = tX + eX; new[2][x + 1] = tY + eY
            rows[d] = (l, h, new[0][x + 1].copy(), new[1][x + 1].copy(), new[2][x + 1].copy(), eM, eY, eX, U)
            F2, F1 = F1, new
            U2, U1 = U1, U
            if d % K == 0:
                pend = measure(new, l, h)
            tb = d >= tracedBackTo + minDiags and (h - l + 1) <= 2 * e + 1
            if d == D or tb:
                break
        Dt = d; dcur = d
        atEnd = d == D
        # ---------------- traceback
        tbf = Dt - (0 if atEnd else tbDiags + 1)
        G1 = blank(); G2 = blank()     # d+1, d+2
        UB1 = 0; UB2 = 0
        ev_ = rendv if atEnd else endv
        total = None; count = 0
        pendB = 0
        for d in range(Dt, tracedBackTo, -1):
            l, h, FM, FX, FY, eM, eY, eX, UF = rows[d]
            x = np.arange(l, h + 1)
            if d == Dt:
                UB = 0
                bM = np.full(len(x), ev_[0], dtype=np.float32) + g * x.astype(np.float32)
                bX = np.full(len(x), ev_[1], dtype=np.float32) + g * x.astype(np.float32)
                bY = np.full(len(x), ev_[2], dtype=np.float32) + g * x.astype(np.float32)
                # units so that values are small: shift by the max
                sh = int(np.rint(float(max(bM.max(), bX.max(), bY.max()))))
                UB = sh; bM = bM - f32(sh); bX = bX - f32(sh); bY = bY - f32(sh)
            else:
                UB = UB1 + pendB; pendB = 0
                s1 = DT(UB1 - UB); s2 = DT(UB2 - UB)
                own = [a[x + 1] for a in G1]; R1 = [a[x + 2] for a in G1]; R2 = [a[x + 2] for a in G2]
                gm2 = R2[0]; gx1 = R1[1]; gy1 = own[2]
                bM = la(la(gm2 + (tMC + s2 - g), gy1 + (tOY + s1)), gx1 + (tOX + s1 - g))
                bX = la(gm2 + (tMX + s2 - g), gx1 + (tEX + s1 - g))
                bY = la(gm2 + (tMY + s2 - g), gy1 + (tEY + s1))
            if d <= tbf:
                if count % 10 == 0:
                    # term 1
                    iF = np.where(np.isfinite(FM), np.rint(FM), f32(0)).astype(np.float32) if precise else np.zeros(len(FM), dtype=np.float32)
                    c1 = la(la((FM - iF) + bM, (FX - iF) + bX), (FY - iF) + bY)
                    acc = -np.inf
                    for v, u in zip(c1, iF):
                        acc = O.log_add(acc, float(v) + float(u))
                    tot = float(acc) + UF + UB
                    tbase = UF + UB
                    if d < Dt:
                        l2_, h2_ = lo[d + 1], hi[d + 1]
                        lm, hm, FMp, FXp, FYp, _, _, _, UFp = rows[d - 1]
                        x2 = np.arange(l2_, h2_ + 1)
                        ok = (x2 - 1 >= lm) & (x2 - 1 <= hm)
                        x2 = x2[ok]
                        if len(x2):
                            ii = x2 - 1 - lm
                            iF = np.where(np.isfinite(FMp[ii]), np.rint(FMp[ii]), f32(0)).astype(np.float32) if precise else np.zeros(len(ii), dtype=np.float32)
                            md = la(la((FMp[ii] - iF) + tMC, (FXp[ii] - iF) + tMX), (FYp[ii] - iF) + tMY)
                            val = (md + G1[0][x2 + 1])
                            acc2 = -np.inf
                            for v, u in zip(val, iF):
                                acc2 = O.log_add(acc2, float(v) + float(u) + ((UFp + UB1 - float(g)) - tbase))
                            acc = O.log_add(acc, acc2)
                            tot = float(acc) + tbase
                    total = tot
                count += 1
                totals[d] = total
                if precise:
                    ci = int(np.floor((UF + UB) - total)); cf = ((UF + UB) - total) - ci
                    lp = (((FM + DT(ci)) + bM) + DT(cf))
                else:
                    lp = (FM + bM) + DT((UF + UB) - total)
                p = np.exp(lp.astype(np.float64))
                okp = (p >= thr) & (x >= 1) & (x <= d - 1)
                for xx, pp in zip(x[okp], p[okp]):
                    pairs.append((int(np.floor(min(pp, 1.0) * 1e7)), int(xx - 1), int(d - xx - 1)))
                if okp.any():
                    stats.setdefault("sumF", 0.0); stats.setdefault("sumB", 0.0); stats.setdefault("n", 0)
                    stats["sumF"] += float(np.abs(FM[okp]).sum()); stats["sumB"] += float(np.abs(bM[okp]).sum()); stats["n"] += int(okp.sum())
                    stats["maxabs_ridgeF"] = max(stats["maxabs_ridgeF"], float(np.abs(FM[okp]).max()))
                    stats["maxabs_ridgeB"] = max(stats["maxabs_ridgeB"], float(np.abs(bM[okp]).max()))
            new = blank()
            if eM is None:
                eM, eY, eX = emissions(d, x) if d > 0 else (NI, NI, NI)
            new[0][x + 1] = bM + eM; new[1][x + 1] = bX + eX; new[2][x + 1] = bY + eY
            G2, G1 = G1, new
            UB2, UB1 = UB1, UB
            if d % K == 0:
                pendB = measure(new, l, h, FM)
        tracedBackTo = tbf
        # restore forward
        if tracedBackTo < D:
            F1 = blank(np.float64 if FWD64 else None); F2 = blank(np.float64 if FWD64 else None)
            l, h, FM, FX, FY, _, _, _, U1 = rows[Dt]
            F1[0][l + 1:h + 2] = FM; F1[1][l + 1:h + 2] = FX; F1[2][l + 1:h + 2] = FY
            l, h, FM, FX, FY, _, _, _, U2 = rows[Dt - 1]
            F2[0][l + 1:h + 2] = FM; F2[1][l + 1:h + 2] = FX; F2[2][l + 1:h + 2] = FY
            pend = 0
    return np.array(pairs, dtype=np.int64).reshape(-1, 3), totals, stats


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lx", type=int, default=1500)
    ap.add_argument("--e", type=int, default=64)
    ap.add_argument("--K", type=int, default=4)
    ap.add_argument("--g", type=float, default=0.0)
    ap.add_argument("--n", type=int, default=2)
    ap.add_argument("--first", type=int, default=100)
    ap.add_argument("--expanded", type=int, default=0)
    ap.add_argument("--cwin", type=int, default=-1)
    ap.add_argument("--ridgeB", type=int, default=0)
    ap.add_argument("--precise", type=int, default=0)
    ap.add_argument("--emis64", type=int, default=0)
    ap.add_argument("--state64", type=int, default=0)
    ap.add_argument("--fwd64", type=int, default=0)
    a = ap.parse_args()
    tables = synth.load_model_file(synth.TEMPLATE_MODEL)
    global DT, FWD64
    FWD64 = bool(a.fwd64)
    if a.state64:
        DT = np.float64
    for i in range(a.n):
        r = synth.make_read(tables[0], a.first + i, lX=a.lx)
        m = O.Model(O.THREE_STATE, tables=tables, scale5=r.scale5)
        want, wtot = O.align_banded(m, r.ref, r.events, r.anchors, params=O.default_params(diagonalExpansion=a.e),
                                    ragged=(1, 1), want_totals=True)
        got, gtot, st = run(r, tables, a.e, a.K, a.g, a.expanded, a.cwin, a.ridgeB, a.precise, a.emis64)
        try:
            s = parity.compare_pairs(got, want)
        except AssertionError as ex:
            s = "FAIL " + str(ex)
        mask = ~np.isnan(wtot)
        dt = np.abs(gtot[mask] - wtot[mask]).max()
        st = {k: round(v, 1) for k, v in st.items()}
        st["meanF"] = round(st.pop("sumF") / st["n"], 1); st["meanB"] = round(st.pop("sumB") / st.pop("n"), 1)
        print("read %d lX=%d lY=%d: %s  max|dtotal|=%.3g (rel %.2g)  %s" % (a.first + i, r.lX, r.lY, s, dt, dt / abs(wtot[mask]).max(), st), flush=True)


if __name__ == "__main__":
    main()
