"""Python restatement of the blocked triangular-solve layout built by k_bc_count / k_bc_fill
(rchol_b200/csrc/rcg_blocked.cu).  TEST INFRASTRUCTURE: used to check the device-built layout field by field and, with
tests/blocked_emulator.py, to exercise the algorithm on the CPU (-m "not gpu")."""
import numpy as np
import scipy.sparse as sp

from blocked_emulator import AHDR, BHDR, WBYTES, RBATCH, FC_WPACK, FC_MINB, FC_TAILB, FC_COLCAP, r16, w_pair_off, rec_batches, fold_batches, fold_bytesA, wp_pair_off


def tree_depths(nb):
    """Depth of every block of the reference's post-order layout (rchol_lap.cpp:254-261)."""
    depth = np.zeros(nb, np.int64)

    def rec(start, total, d):
        if total == 1:
            depth[start] = d
        else:
            depth[start + total - 1] = d
            h = (total - 1) // 2
            rec(start, h, d + 1)
            rec(start + h, h, d + 1)
    rec(0, nb, 0)
    return depth


def direction_matrix(G, part, backward):
    """Lower-triangular solve matrix of a direction in its solve index space + block bounds / depths."""
    rp, ci, v = (np.asarray(a) for a in G)
    N = len(rp) - 1
    U = sp.csr_matrix((v, ci.astype(np.int64), rp.astype(np.int64)), shape=(N, N))
    bounds = np.array([0, N], np.int64) if part is None or len(part) < 2 else np.asarray(part, np.int64)
    nb = len(bounds) - 1
    depth = tree_depths(nb)
    if not backward:
        L = U.T.tocsr()
    else:
        J = np.arange(N)[::-1]
        L = U[J][:, J].tocsr()
        bounds = N - bounds[::-1]
        depth = depth[::-1]
    L.sort_indices()
    return L, bounds, depth


def build_layout(L, bounds, depth, root_first, Kr=2, E=16, Dfar=128, Dfar_sep=32, reversed_=False, tile_sep=1, tile_leaf=8, E_sep=None,
                 fold=False, wb_min=0, Dfar_wb=64, wb_jagged=True):
    N = L.shape[0]
    rp, col, val = L.indptr.astype(np.int64), L.indices.astype(np.int64), L.data
    nb = len(bounds) - 1
    max_depth = int(np.max(depth)) if nb else 0
    Dfar_sep = min(Dfar, Dfar_sep)
    dfar_of = [Dfar if depth[b] == max_depth else Dfar_sep for b in range(nb)]
    tile_of = [tile_leaf if depth[b] == max_depth else tile_sep for b in range(nb)]   # chunks per far tile
    E_sep = min(E, 6) if E_sep is None else E_sep
    e_of = [E if depth[b] == max_depth else E_sep for b in range(nb)]                 # early/late distance
    kr_of = [Kr] * nb
    # warp-per-block levels (k_wb_solve): tree levels with at least wb_min non-empty blocks: nothing folded, one jagged class
    per_depth = {}
    for b in range(nb):
        if bounds[b + 1] > bounds[b]:
            per_depth[int(depth[b])] = per_depth.get(int(depth[b]), 0) + 1
    wb_of = [1 if wb_min > 0 and per_depth.get(int(depth[b]), 0) >= wb_min else 0 for b in range(nb)]
    nch_depth = {}
    for b in range(nb):
        if bounds[b + 1] > bounds[b]:
            nch_depth[int(depth[b])] = max(nch_depth.get(int(depth[b]), 0), (int(bounds[b + 1] - bounds[b]) + 31) // 32)
    for b in range(nb):
        if wb_of[b]:
            # window of the level: leaves two planes of a 3-D box (rows^(2/3) each, at most Dfar_wb chunks), separator
            # blocks the whole block (at most 128 chunks)
            want, cap = nch_depth[int(depth[b])], 128
            if depth[b] == max_depth:
                want, cap = int(2.0 * (32.0 * nch_depth[int(depth[b])]) ** (2.0 / 3.0) / 32.0) + 1, Dfar_wb
            dw = 4
            while dw < want and dw < cap:
                dw <<= 1
            kr_of[b], e_of[b], dfar_of[b], tile_of[b] = 0, (0 if wb_jagged else dw), dw, tile_sep
    Dfar_leaf = Dfar
    chunk0 = np.zeros(nb + 1, np.int64)
    tile0 = np.zeros(nb + 1, np.int64)
    for b in range(nb):
        nch = (bounds[b + 1] - bounds[b] + 31) // 32
        chunk0[b + 1] = chunk0[b] + nch
        tile0[b + 1] = tile0[b] + (nch + tile_of[b] - 1) // tile_of[b]
    nchunks, ntiles = int(chunk0[nb]), int(tile0[nb])
    blobsA, blobsB = [None] * nchunks, [None] * nchunks
    tile_need = np.zeros(ntiles, np.uint32)
    far_rows = [None] * N
    for b in range(nb):
        blo, bhi = int(bounds[b]), int(bounds[b + 1])
        Dfar = dfar_of[b]
        wmask = 32 * Dfar - 1
        for k in range((bhi - blo + 31) // 32):
            g = int(chunk0[b]) + k
            rows = [j for j in range(blo + 32 * k, min(bhi, blo + 32 * k + 32))]
            nr = len(rows)
            c_far = blo + 32 * max(0, k + 1 - Dfar)
            c_early = blo + 32 * max(0, k - e_of[b])
            Dk = kr_of[b]
            if fold and not wb_of[b]:
                # fold depth of the chunk (chunk_fold_depth): the largest D <= Kr with at most FC_COLCAP distinct columns in
                # the D previous chunks; the previous chunk is always folded
                ng = min(k, kr_of[b])
                by_dist = [set() for _ in range(ng + 1)]
                for j in rows:
                    s_, e_ = rp[j], rp[j + 1]
                    for c in col[s_:e_ - 1]:
                        if blo + 32 * (k - ng) <= c < blo + 32 * k:
                            by_dist[k - (int(c) - blo) // 32].add(int(c))
                Dk, cum = 0, 0
                for dd in range(1, ng + 1):
                    if dd > 1 and cum + len(by_dist[dd]) > FC_COLCAP:
                        break
                    cum += len(by_dist[dd])
                    Dk = dd
            c_late = blo + 32 * max(0, k - Dk)
            c_rec = blo + 32 * k
            parts = []
            D = np.zeros((32, 32))
            for l, j in enumerate(rows):
                s, e = rp[j], rp[j + 1]
                assert col[e - 1] == j
                cj, vj = col[s:e - 1], val[s:e - 1]
                m_far = cj < c_far
                m_early = (cj >= c_far) & (cj < c_early)
                m_late = (cj >= c_early) & (cj < c_late)
                m_rec = (cj >= c_late) & (cj < c_rec)
                m_diag = cj >= c_rec
                parts.append((cj[m_early], vj[m_early], cj[m_late], vj[m_late], cj[m_rec], vj[m_rec]))
                D[l, cj[m_diag] - c_rec] = vj[m_diag]
                D[l, l] = val[e - 1]
                fc = cj[m_far]
                far_rows[j] = ((N - 1 - fc) if reversed_ else fc, vj[m_far])
                loc = fc[fc >= blo]
                if len(loc):
                    t = int(tile0[b]) + k // tile_of[b]
                    tile_need[t] = max(tile_need[t], (int(loc[-1]) - blo) // 32 + 1)
            for l in range(nr, 32):
                D[l, l] = 1.0
                parts.append((np.zeros(0, np.int64), np.zeros(0), np.zeros(0, np.int64), np.zeros(0), np.zeros(0, np.int64), np.zeros(0)))
            W = np.zeros((32, 32))
            for i in range(32):
                e_i = np.zeros(32)
                e_i[i] = 1.0
                W[i] = (e_i - D[i, :i] @ W[:i]) / D[i, i]
            n_early = np.array([len(p[0]) for p in parts])
            n_late = np.array([len(p[2]) for p in parts])
            n_rec = np.array([len(p[4]) for p in parts])
            nslots, nl, ne_max, ne_tot = int(n_rec.max()), int(n_late.max()), int(n_early.max()), int(n_early.sum())
            # blob A
            if wb_of[b]:
                a = np.zeros(16, np.uint8)
                a[:12].view(np.uint32)[:] = [0, nr, 0]
                blobsA[g] = a
            elif fold:
                # dense panel M = Winv L_rec over the distinct recent columns (ascending)
                cols = sorted(set(int(c) for p in parts for c in p[4]))
                ncol = len(cols)
                ncb = fold_batches(ncol)
                pos = {c: i for i, c in enumerate(cols)}
                Lrec = np.zeros((32, ncol))
                for l, p in enumerate(parts):
                    for c, v in zip(p[4], p[5]):
                        Lrec[l, pos[int(c)]] = v
                M = W @ Lrec
                a = np.zeros(fold_bytesA(ncb), np.uint8)
                a[:12].view(np.uint32)[:] = [ncb, nr, ncol]
                nbody, npad = ncb - FC_MINB, 4 * ncb - ncol
                offs = np.full(4 * ncb, 8 * (wmask + 1), np.uint32)         # column slots, padding first
                offs[npad:] = [8 * ((c - blo) & wmask) for c in cols]
                vals = np.zeros((ncb, 2, 32, 2))
                for i in range(ncol):
                    ci = npad + i
                    vals[ci >> 2, (ci >> 1) & 1, :, ci & 1] = M[:, i]
                # blob: header | tail offsets | tail values | body offsets | body values
                a[16:16 + 16 * FC_MINB] = offs[4 * nbody:].view(np.uint8)
                a[16 + 16 * FC_MINB:FC_TAILB] = vals[nbody:].reshape(-1).view(np.uint8)
                a[FC_TAILB:FC_TAILB + 16 * nbody] = offs[:4 * nbody].view(np.uint8)
                a[FC_TAILB + 16 * nbody:] = vals[:nbody].reshape(-1).view(np.uint8)
                blobsA[g] = a
            nbt = rec_batches(nslots)
            a = np.zeros(AHDR + WBYTES + RBATCH * nbt, np.uint8)
            a[:12].view(np.uint32)[:] = [nbt, nr, nslots]
            wp = a[AHDR:AHDR + WBYTES].view(np.float64)
            for pp in range(16):
                for row in range(32):
                    o = w_pair_off(pp, row) // 8
                    wp[o] = W[row, 2 * pp]
                    wp[o + 1] = W[row, 2 * pp + 1]
            for bt in range(nbt):
                R = a[AHDR + WBYTES + RBATCH * bt: AHDR + WBYTES + RBATCH * (bt + 1)]
                vals = R[:2048].view(np.float64).reshape(32, 4, 2)      # [row][pair][2]
                offs = R[2048:].view(np.uint32).reshape(32, 2, 4)        # [row][half][4]
                offs[:] = 8 * (wmask + 1)
                for l, p in enumerate(parts):
                    for u in range(8):
                        sidx = 8 * bt + u
                        if sidx < len(p[4]):
                            vals[l, u >> 1, u & 1] = p[5][sidx]
                            offs[l, u >> 2, u & 3] = 8 * ((p[4][sidx] - blo) & wmask)
            if not fold and not wb_of[b]:
                blobsA[g] = a
            # blob B
            order = sorted(range(32), key=lambda l: (-n_early[l], l))
            rank = np.zeros(32, np.int64)
            rank[order] = np.arange(32)
            bb = np.zeros(BHDR + r16(ne_max) + r16(8 * ne_tot) + r16(2 * ne_tot) + 320 * nl + (FC_WPACK if (fold or wb_of[b]) else 0), np.uint8)
            bb[:16].view(np.uint32)[:] = [ne_max, ne_tot, nl, Dk]
            bb[16:48] = order
            bb[48:80] = rank
            o = BHDR + r16(ne_max)
            ev = bb[o: o + 8 * ne_tot].view(np.float64)
            o += r16(8 * ne_tot)
            ec = bb[o: o + 2 * ne_tot].view(np.uint16)
            o += r16(2 * ne_tot)
            lv = bb[o: o + 256 * nl].view(np.float64).reshape(nl, 32)
            lc = bb[o + 256 * nl: o + 320 * nl].view(np.uint16).reshape(nl, 32)
            if fold or wb_of[b]:
                wq = bb[o + 320 * nl:].view(np.float64)
                for pp in range(16):
                    for row in range(2 * pp, 32):
                        wq[wp_pair_off(pp, row) // 8] = W[row, 2 * pp]
                        wq[wp_pair_off(pp, row) // 8 + 1] = W[row, 2 * pp + 1]
            base = 0
            for s in range(ne_max):
                cs = int((n_early > s).sum())
                bb[BHDR + s] = cs
                for l in range(32):
                    if n_early[l] > s:
                        ev[base + rank[l]] = parts[l][1][s]
                        ec[base + rank[l]] = (parts[l][0][s] - blo) & wmask
                base += cs
            lc[:] = wmask + 1
            for l, p in enumerate(parts):
                lv[:len(p[2]), l] = p[3]
                lc[:len(p[2]), l] = (p[2] - blo) & wmask
            blobsB[g] = bb
    offA = np.concatenate([[0], np.cumsum([len(a) for a in blobsA])]).astype(np.int64)
    offB = np.concatenate([[0], np.cumsum([len(a) for a in blobsB])]).astype(np.int64)
    far_rp = np.concatenate([[0], np.cumsum([len(r[0]) for r in far_rows])]).astype(np.int64)
    blocks = []
    for gl in range(max_depth + 1):
        want = gl if root_first else max_depth - gl
        level = [b for b in range(nb) if depth[b] == want and bounds[b + 1] > bounds[b]]
        if level and wb_of[level[0]]:     # warp-per-block level: longest block first (dynamic hand-out), stable
            level.sort(key=lambda b: -(int(bounds[b + 1]) - int(bounds[b])))
        for b in level:
            blocks.append([bounds[b], bounds[b + 1], chunk0[b], tile0[b], len(blocks), dfar_of[b], tile_of[b],
                           e_of[b] | (kr_of[b] << 8) | (wb_of[b] << 16)])
    cat = lambda xs, dt: np.concatenate(xs).astype(dt) if len(xs) else np.zeros(0, dt)
    return dict(fold=int(fold), active=1, nchunks=nchunks, ntiles=ntiles, nblocks=len(blocks), N=N, Kr=Kr, E=E, Dfar=Dfar_leaf, Dfar_sep=Dfar_sep,
                offA=offA, offB=offB, blobA=cat(blobsA, np.uint8), blobB=cat(blobsB, np.uint8), far_rp=far_rp,
                far_col=cat([r[0] for r in far_rows], np.uint32), far_val=cat([r[1] for r in far_rows], np.float64),
                tile_need=tile_need, blocks=np.array(blocks, np.uint32).reshape(-1, 8))


def compare_layouts(dev, ref, val_tol=1e-12):
    """Device-built layout against the Python restatement: integers exact, values to val_tol (relative to max)."""
    for k in ("nchunks", "ntiles", "nblocks", "N", "Kr", "E", "Dfar", "Dfar_sep"):
        assert dev[k] == ref[k], k
    for k in ("offA", "offB", "far_rp", "far_col", "tile_need", "blocks"):
        assert np.array_equal(dev[k], ref[k]), k
    np.testing.assert_allclose(dev["far_val"], ref["far_val"], rtol=0, atol=0)
    for name, off in (("blobA", "offA"), ("blobB", "offB")):
        da, ra = dev[name], ref[name]
        assert len(da) == len(ra), name
    # blob A: header + values + column slots
    for g in range(ref["nchunks"]):
        a, r = dev["blobA"][ref["offA"][g]: ref["offA"][g + 1]], ref["blobA"][ref["offA"][g]: ref["offA"][g + 1]]
        assert np.array_equal(a[:12], r[:12]), ("A header", g)
        if len(r) == 16:
            assert len(a) == 16
        elif ref.get("fold"):
            ncb = int(r[:12].view(np.uint32)[0])
            nbody = ncb - FC_MINB
            for lo_, hi_ in ((16, 16 + 16 * FC_MINB), (FC_TAILB, FC_TAILB + 16 * nbody)):
                assert np.array_equal(a[lo_:hi_], r[lo_:hi_]), ("panel columns", g)
            for lo_, hi_ in ((16 + 16 * FC_MINB, FC_TAILB), (FC_TAILB + 16 * nbody, len(r))):
                md, mr = a[lo_:hi_].view(np.float64), r[lo_:hi_].view(np.float64)
                if len(mr):
                    assert np.abs(md - mr).max() <= val_tol * max(1.0, np.abs(mr).max()), ("panel", g)
        else:
            nslots = int(r[:12].view(np.uint32)[2])
            wd, wr = a[AHDR:AHDR + WBYTES].view(np.float64), r[AHDR:AHDR + WBYTES].view(np.float64)
            assert np.abs(wd - wr).max() <= val_tol * max(1.0, np.abs(wr).max()), ("Winv", g)
            assert np.array_equal(a[AHDR + WBYTES:], r[AHDR + WBYTES:]), ("recent", g, nslots)
        b, rb = dev["blobB"][ref["offB"][g]: ref["offB"][g + 1]], ref["blobB"][ref["offB"][g]: ref["offB"][g + 1]]
        assert np.array_equal(b[:12], rb[:12]) and np.array_equal(b[16:80], rb[16:80]), ("B header", g)
        ne_max, ne_tot, nl = (int(v) for v in rb[:12].view(np.uint32))
        assert np.array_equal(b[BHDR:BHDR + ne_max], rb[BHDR:BHDR + ne_max]), ("cnt", g)
        o = BHDR + r16(ne_max)
        assert np.array_equal(b[o:o + 8 * ne_tot], rb[o:o + 8 * ne_tot]), ("early val", g)
        o += r16(8 * ne_tot)
        assert np.array_equal(b[o:o + 2 * ne_tot], rb[o:o + 2 * ne_tot]), ("early col", g)
        o += r16(2 * ne_tot)
        assert np.array_equal(b[o:o + 320 * nl], rb[o:o + 320 * nl]), ("late", g)
        if len(rb) > o + 320 * nl:
            wd, wr = b[o + 320 * nl:].view(np.float64), rb[o + 320 * nl:].view(np.float64)
            assert len(wd) == len(wr) == FC_WPACK // 8
            assert np.abs(wd - wr).max() <= val_tol * max(1.0, np.abs(wr).max()), ("Winv packed", g)
