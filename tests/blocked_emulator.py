"""Host interpreter of the blocked triangular-solve layout (TEST INFRASTRUCTURE, never on the product path).

``solve_from_layout`` replays, chunk by chunk and with the same shared-memory window ring, what k_bc_solve
(rchol_b200/csrc/rcg_blocked.cu) does with the arrays that the set-up kernels built on the device.  It separates
"the set-up produced a wrong layout" from "the solve kernel mis-reads a right layout" in one GPU run.
"""
import numpy as np

AHDR, BHDR, WBYTES, RBATCH = 16, 80, 1024 * 8, 3072
FC_MINB, FC_WPACK = 4, 4352          # folded layout (chain_mode 5, rcg_fold.cuh)
FC_COLCAP = 32                       # panel columns per chunk beyond the previous chunk's (chunk_fold_depth)


FC_TAILB = 16 + 1040 * FC_MINB       # header + tail (the newest 16 columns): byte offset of the body


def fold_batches(ncol):
    """Tail (4 batches) + an even number of body batches."""
    return FC_MINB if ncol <= 4 * FC_MINB else FC_MINB + 2 * ((ncol - 4 * FC_MINB + 7) // 8)


def fold_bytesA(ncb):
    return 16 + 1040 * ncb


def wp_pair_off(p, row):
    """Byte offset of {Winv[row][2p], Winv[row][2p+1]} (row >= 2p) in the packed lower triangle of blob B."""
    return 16 * (p * (33 - p) + row - 2 * p)


def unpack_winv_packed(raw_bytes):
    W = np.zeros((32, 32))
    d = raw_bytes.view(np.float64)
    for p in range(16):
        for row in range(2 * p, 32):
            o = wp_pair_off(p, row) // 8
            W[row, 2 * p], W[row, 2 * p + 1] = d[o], d[o + 1]
    return W


def unpack_panel(a):
    """Folded blob A -> (ncb, nr, ncol, window slots [4*ncb], panel values [32, 4*ncb])."""
    ncb, nr, ncol = (int(v) for v in a[:12].view(np.uint32))
    if ncb == 0:                      # block of a warp-per-block level: bare header, nothing folded
        assert len(a) == 16 and ncol == 0
        return 0, nr, 0, np.zeros(0, np.int64), np.zeros((32, 0))
    assert len(a) == fold_bytesA(ncb) and ncb == fold_batches(ncol)
    nbody = ncb - FC_MINB
    # column order: body batches first (oldest, padding in front), then the tail; the tail sits first in the blob
    offs = np.concatenate([a[FC_TAILB:FC_TAILB + 16 * nbody], a[16:16 + 16 * FC_MINB]]).view(np.uint32).astype(np.int64)
    assert np.all(offs % 8 == 0)
    vals = np.concatenate([a[FC_TAILB + 16 * nbody:], a[16 + 16 * FC_MINB:FC_TAILB]]).view(np.float64).reshape(ncb, 2, 32, 2)
    M = vals.transpose(2, 0, 1, 3).reshape(32, 4 * ncb)                    # [batch][pair][row][2] -> [row][column]
    return ncb, nr, ncol, offs // 8, M


def r16(v):
    return (v + 15) & ~15


def w_pair_off(p, row):
    """Byte offset of {Winv[row][2p], Winv[row][2p+1]} in the W part of blob A (rcg_blocked.cu)."""
    return 512 * p + 16 * row


def rec_batches(nslots):
    return 1 if nslots <= 8 else (nslots + 7) // 8


def unpack_winv(raw):
    """1024 doubles [column pair][row] double2 -> dense 32x32."""
    return raw.reshape(16, 32, 2).transpose(1, 0, 2).reshape(32, 32).copy()


def unpack_recent(raw_bytes, nbt):
    """Batches of 8 recent slots -> (values [8*nbt, 32], window slots [8*nbt, 32])."""
    rv = np.zeros((8 * nbt, 32))
    rc = np.zeros((8 * nbt, 32), np.int64)
    for bt in range(nbt):
        R = raw_bytes[RBATCH * bt: RBATCH * (bt + 1)]
        vals = R[:2048].view(np.float64).reshape(32, 4, 2)      # [row][pair][2]
        offs = R[2048:].view(np.uint32).reshape(32, 2, 4)        # [row][half][4] byte offsets into the window
        for u in range(8):
            rv[8 * bt + u] = vals[:, u >> 1, u & 1]
            rc[8 * bt + u] = offs[:, u >> 2, u & 3] // 8
    return rv, rc


def solve_from_layout(lay, rhs, reversed_):
    N = lay["N"]
    out = np.zeros(N)
    A, B = lay["blobA"], lay["blobB"]
    far_rp, far_col, far_val = lay["far_rp"], lay["far_col"].astype(np.int64), lay["far_val"]
    fold = bool(lay.get("fold", 0))
    stats = dict(chunks=0, rec_slots=0, late_slots=0, early_max=0, early_tot=0, panel_cols=0)
    for lo, hi, chunk0, tile0, gidx, dfar, tile, *_ in lay["blocks"].astype(np.int64):
        nch = (hi - lo + 31) // 32
        wrows = 32 * dfar          # window of this block (leaves: Dfar, separators: Dfar_sep)
        wmask = wrows - 1
        win = np.full(wrows + 1, np.nan)
        win[wrows] = 0.0   # the slot padding entries point at
        for k in range(nch):
            g = chunk0 + k
            j = lo + 32 * k + np.arange(32)
            valid = j < hi
            t0 = np.zeros(32)
            for l in np.nonzero(valid)[0]:
                jj = j[l]
                i = N - 1 - jj if reversed_ else jj
                s, e = far_rp[jj], far_rp[jj + 1]
                t0[l] = rhs[i] - np.dot(far_val[s:e], out[far_col[s:e]])
            need = lay["tile_need"][tile0 + k // tile]
            assert need <= tile * (k // tile), "far tile would wait for a chunk of its own tile"
            # ---- blob B: early (jagged diagonals) + late (ELL) --------------------------------------------
            b = B[lay["offB"][g]: lay["offB"][g + 1]]
            ne_max, ne_tot, nl = (int(v) for v in b[:12].view(np.uint32))
            perm, rank = b[16:48].astype(np.int64), b[48:80].astype(np.int64)
            assert sorted(perm) == list(range(32)) and np.array_equal(perm[rank], np.arange(32))
            cnt = b[BHDR: BHDR + ne_max].astype(np.int64)
            o = BHDR + r16(ne_max)
            ev = b[o: o + 8 * ne_tot].view(np.float64)
            o += r16(8 * ne_tot)
            ec = b[o: o + 2 * ne_tot].view(np.uint16).astype(np.int64)
            o += r16(2 * ne_tot)
            lv = b[o: o + 256 * nl].view(np.float64).reshape(nl, 32)
            lc = b[o + 256 * nl: o + 320 * nl].view(np.uint16).astype(np.int64).reshape(nl, 32)
            a = A[lay["offA"][g]: lay["offA"][g + 1]]
            wb_chunk = len(a) == 16           # block of a warp-per-block level: bare header in blob A, Winv packed in blob B
            assert o + 320 * nl + (FC_WPACK if (fold or wb_chunk) else 0) == len(b) and cnt.sum() == ne_tot
            ts = t0[perm]
            base = 0
            for s in range(ne_max):
                cs = cnt[s]
                ts[:cs] -= ev[base: base + cs] * win[ec[base: base + cs]]
                base += cs
            t = ts[rank]
            for s in range(nl):
                t -= lv[s] * win[lc[s]]
            if fold or wb_chunk:
                # ---- helper: u = Winv t ; chain: x = u - M x_rec (dense panel, blob A) ---------------------
                W = unpack_winv_packed(b[o + 320 * nl:])
                ncb, nr, ncol, slots, M = unpack_panel(a)
                assert nr == int(valid.sum())
                npad = 4 * ncb - ncol
                assert np.all(slots[:npad] == wrows) and not M[:, :npad].any(), "padding columns come first"
                x = W @ t - M @ win[slots]
                win[(32 * k + np.arange(32)) & wmask] = x
                jv = j[valid]
                out[(N - 1 - jv) if reversed_ else jv] = x[valid]
                stats["chunks"] += 1
                stats["panel_cols"] += ncol
                stats["late_slots"] += nl
                stats["early_max"] += ne_max
                stats["early_tot"] += ne_tot
                continue
            # ---- blob A: Winv + recent (ELL) ------------------------------------------------------------
            nbt, nr, nslots = (int(v) for v in a[:12].view(np.uint32))
            assert nr == int(valid.sum()) and nbt == rec_batches(nslots) and len(a) == AHDR + WBYTES + RBATCH * nbt
            W = unpack_winv(a[AHDR: AHDR + WBYTES].view(np.float64))
            rv, rc = unpack_recent(a[AHDR + WBYTES:], nbt)
            for s in range(8 * nbt):
                t -= rv[s] * win[rc[s]]
            x = W @ t
            win[(32 * k + np.arange(32)) & wmask] = x
            jv = j[valid]
            out[(N - 1 - jv) if reversed_ else jv] = x[valid]
            stats["chunks"] += 1
            stats["rec_slots"] += nslots
            stats["late_slots"] += nl
            stats["early_max"] += ne_max
            stats["early_tot"] += ne_tot
    return out, stats
