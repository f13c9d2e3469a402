"""CPU ORACLE (test infrastructure, NOT product code) -- numpy restatement of the reference's
polar hot path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; nothing under polar-code-pytorch-sionna_b200/ does.

Parity status: PINNED.  Every function below is checked by tests/test_oracle_golden.py against
fixtures under tests/golden/ that were produced by importing the unmodified reference in the build
container (oracle/gen_golden.py, committed), plus the reference's only published result
(README-config BLER curve, x_run_sn_polar/plots/sc_c.mc_iter=1_c.bs=100.png).

All file:line citations are relative to the reference checkout (/root/reference).

Conventions
  logits : ln P(1)/P(0), fp32, [B, n]   (decoder input, polar_sc.py:113-122)
  llr    : -logits ("true" LLR ln P0/P1)  (polar_sc.py:122, polar_scl.py:219)
  frozen : uint8[n], 1 = frozen position  (polar_sc.py:23-24)
  tree   : natural (non bit-reversed) order; node (a, len) has left [a, a+len/2), right [a+len/2, a+len)
"""
import numpy as np

LLR_MAX = 30.0  # polar_sc.py:21, polar_scl.py:35


# ------------------------------------------------------------------------------------------------
# A1  frozen-set construction            x_run_sn_polar/polar/froze.py:4-16
# ------------------------------------------------------------------------------------------------
def arikan_G(n):
    """G = F2^{(x) log2 n} with F2 = [[1,0],[1,1]] (d_kernels.py:8-9, froze.py:9-12).
    Closed form: G[i, j] = 1  <=>  (i & j) == j   (row i has 2^popcount(i) ones)."""
    i = np.arange(n)[:, None]
    j = np.arange(n)[None, :]
    return ((i & j) == j).astype(np.float32)


def rm_frozen_pos(n, f_num):
    """froze.py:13-14: frozen_pos = sort(argsort(row_weights)[:f_num]).
    The reference uses torch.argsort (unstable) on a CPU fp32 tensor; ties at the cut are resolved by
    that op, so the oracle calls the very same op on the very same values (SURVEY A1 hazard)."""
    import torch
    w = torch.from_numpy(arikan_G(n)).sum(dim=1)
    return torch.sort(torch.argsort(w)[:f_num])[0].numpy().astype(np.int64)


def frozen_vec(frozen_pos, n):
    f = np.zeros(n, dtype=np.uint8)
    f[np.asarray(frozen_pos, dtype=np.int64)] = 1
    return f


def info_positions(frozen_pos, n):
    """polar_sc.py:19 / enc.py:24: np.setdiff1d(arange(n), frozen_pos) (ascending)."""
    return np.setdiff1d(np.arange(n), np.asarray(frozen_pos)).astype(np.int64)


# ------------------------------------------------------------------------------------------------
# A2/A3  encoder            x_run_sn_polar/polar/enc.py:30-43, my_sn/fec/polar/enc.py:85-113
# ------------------------------------------------------------------------------------------------
def polar_transform(u_full):
    """x = u.G over GF(2):  c_j = XOR_{i superset j} u_i.  Stage s pairs d with d+2^s for every d
    whose bit s is clear (my_sn/fec/polar/enc.py:70-74,89-92).  uint8 [B, n] -> uint8 [B, n]."""
    x = np.array(u_full, dtype=np.uint8, copy=True)
    n = x.shape[-1]
    s = 1
    while s < n:
        v = x.reshape(x.shape[0], n // (2 * s), 2, s)
        v[:, :, 0, :] ^= v[:, :, 1, :]
        s *= 2
    return x


def encode(u, frozen_pos, n):
    """enc.py:33-42: scatter the k info bits to info_pos (ascending), frozen = 0, multiply by G mod 2."""
    u = np.asarray(u)
    info = info_positions(frozen_pos, n)
    full = np.zeros((u.shape[0], n), dtype=np.uint8)
    full[:, info] = (u != 0).astype(np.uint8)
    return polar_transform(full)


# ------------------------------------------------------------------------------------------------
# A4-A7  SC decoder (fp32, min-sum)       x_run_sn_polar/polar/polar_sc.py:33-133
# ------------------------------------------------------------------------------------------------
def _f32(a, b):
    """polar_sc.py:33-48: clip both to +-30, then sign.sign.min|.| (line 46 overrides the boxplus)."""
    a = np.clip(a, -LLR_MAX, LLR_MAX).astype(np.float32)
    b = np.clip(b, -LLR_MAX, LLR_MAX).astype(np.float32)
    return (np.sign(a) * np.sign(b) * np.minimum(np.abs(a), np.abs(b))).astype(np.float32)


def _g32(a, b, u):
    """polar_sc.py:49-53: (1-2u).a + b, unclipped, fp32 (product with +-1 is exact: one rounding)."""
    return ((np.float32(1) - np.float32(2) * u.astype(np.float32)) * a + b).astype(np.float32)


def sc_decode_full(logits, frozen):
    """Returns u_hat for ALL n positions, uint8 [B, n] (msg_uhat[:,0,:], polar_sc.py:112).
    Recursion restated from polar_sc.py:54-98 (depth first: f -> left -> g -> right -> combine)."""
    llr = (np.float32(-1.0) * np.asarray(logits, dtype=np.float32))  # polar_sc.py:117,122
    B, n = llr.shape
    u = np.zeros((B, n), dtype=np.uint8)

    def rec(a, L):
        """L: [B, len] LLRs of node starting at leaf a; returns partial sums beta uint8 [B, len]."""
        ln = L.shape[1]
        if ln == 1:
            if frozen[a]:
                return np.zeros((B, 1), dtype=np.uint8)            # polar_sc.py:91-92
            bit = (L[:, 0] <= 0).astype(np.uint8)                   # polar_sc.py:94-97 (0 -> 1)
            u[:, a] = bit
            return bit[:, None]
        h = ln // 2
        bl = rec(a, _f32(L[:, :h], L[:, h:]))                       # polar_sc.py:66-68
        br = rec(a + h, _g32(L[:, :h], L[:, h:], bl))              # polar_sc.py:74-81
        return np.concatenate([bl ^ br, br], axis=1)               # polar_sc.py:83-89

    rec(0, llr)
    return u


def _f32_boxplus(a, b):
    """my_sn/fec/polar/dec.py:33-46: clip to +-30, ln(1+e^(x+y)) - ln(e^x+e^y), fp32 like the reference's torch ops."""
    a = np.clip(a, -LLR_MAX, LLR_MAX).astype(np.float32)
    b = np.clip(b, -LLR_MAX, LLR_MAX).astype(np.float32)
    out = np.log(np.float32(1) + np.exp(a + b)).astype(np.float32)
    out -= np.log(np.exp(a) + np.exp(b)).astype(np.float32)
    return out.astype(np.float32)


def sc_decode_boxplus_full(logits, frozen):
    """Sionna-style SC (my_sn/fec/polar/dec.py:47-157): sc_decode_full with the exact boxplus f.  SECONDARY oracle:
    numpy's exp/log are not the reference's (torch) nor CUDA's, so it is compared statistically, never bit-exactly."""
    llr = (np.float32(-1.0) * np.asarray(logits, dtype=np.float32))
    B, n = llr.shape
    u = np.zeros((B, n), dtype=np.uint8)

    def rec(a, L):
        ln = L.shape[1]
        if ln == 1:
            if frozen[a]:
                return np.zeros((B, 1), dtype=np.uint8)
            bit = (L[:, 0] <= 0).astype(np.uint8)
            u[:, a] = bit
            return bit[:, None]
        h = ln // 2
        bl = rec(a, _f32_boxplus(L[:, :h], L[:, h:]))
        br = rec(a + h, _g32(L[:, :h], L[:, h:], bl))
        return np.concatenate([bl ^ br, br], axis=1)

    rec(0, llr)
    return u


def sc_decode(logits, frozen_pos, n):
    """SC_Dec.forward (polar_sc.py:113-133): [B, n] logits -> [B, k] fp32 bits at ascending info_pos."""
    fz = frozen_vec(frozen_pos, n)
    u = sc_decode_full(np.asarray(logits).reshape(-1, n), fz)
    return u[:, info_positions(frozen_pos, n)].astype(np.float32)


# ------------------------------------------------------------------------------------------------
# A8-A12  SCL decoder (fp64, min-sum, exact softplus PM)   x_run_sn_polar/polar/polar_scl.py:49-234
# ------------------------------------------------------------------------------------------------
def _f64(a, b):
    """polar_scl.py:93-106: clip via np.minimum/np.maximum, min-sum (torch ops on fp64)."""
    a = np.maximum(np.minimum(a, LLR_MAX), -LLR_MAX)
    b = np.maximum(np.minimum(b, LLR_MAX), -LLR_MAX)
    return np.sign(a) * np.sign(b) * np.minimum(np.abs(a), np.abs(b))


def _f64_boxplus(a, b, jitter=None):
    """my_sn/fec/polar/dec.py:331-340 (_cn_op_np): clip to +-30, log(1+exp(x+y)) - log(exp(x)+exp(y)), same operation order.
    `jitter`: see _softplus_neg -- every exp / log result moved by -2 .. +2 ulp at random (numpy vs CUDA vs glibc)."""
    x = np.maximum(np.minimum(a, LLR_MAX), -LLR_MAX)
    y = np.maximum(np.minimum(b, LLR_MAX), -LLR_MAX)

    def j(v):
        if jitter is None:
            return v
        d = jitter.integers(-2, 3, size=np.shape(v))
        for _ in range(2):
            v = np.where(d > 0, np.nextafter(v, np.inf), np.where(d < 0, np.nextafter(v, -np.inf), v))
            d = d - np.sign(d)
        return v
    out = j(np.log(1 + j(np.exp(x + y))))
    out = out - j(np.log(j(np.exp(x)) + j(np.exp(y))))
    return out


def _g64(a, b, u):
    """polar_scl.py:107-108."""
    return np.multiply((1 - 2 * u), a) + b


def _softplus_neg(x, use_log1p=False, jitter=None):
    """polar_scl.py:82-83: log(1 + exp(-x)), literal (NOT log1p, NOT max(0,-x)).
    `jitter` (a numpy Generator) moves every result by -1 / 0 / +1 ulp at random: the spread of exp/log
    implementations (numpy SIMD, glibc, CUDA) in the last bit, used to detect codewords whose ranking the
    reference itself does not reproduce across math libraries."""
    v = np.log1p(np.exp(-x)) if use_log1p else np.log(1 + np.exp(-x))
    if jitter is not None:
        d = jitter.integers(-1, 2, size=np.shape(v))
        v = np.where(d > 0, np.nextafter(v, np.inf), np.where(d < 0, np.nextafter(v, -np.inf), v))
    return v


def scl_decode_full(logits, frozen, list_size, use_log1p=False, stable_sort=True, ulp_jitter_seed=None, boxplus=False,
                    fast_nodes=0):
    """Equivalent L-survivor formulation of polar_scl.py:121-209 (bit-exact incl. PMs):
    L paths with pm = [0, 30, ..., 30] (polar_scl.py:192-194; the reference's 2L slots are these L
    paths duplicated pairwise); frozen leaf: pm += softplus(-llr) (u=0); info leaf: fork every path
    (u=0: pm+softplus(-llr), u=1: pm+softplus(+llr)), sort the 2L candidates ascending, keep L.
    Returns (u_hat uint8 [B, L, n] sorted by pm ascending, pm float64 [B, L]).
    `use_log1p` / `stable_sort` / `ulp_jitter_seed` give the oracle *variants* used to detect ill-conditioned
    lists."""
    jit = None if ulp_jitter_seed is None else np.random.default_rng(ulp_jitter_seed)
    # boxplus=True: the Sionna-style list decoder my_sn/fec/polar/dec.py:158-537 with use_fast_scl=False (same recursion,
    # exact boxplus check node); SECONDARY oracle (statistical parity: its exp / log are the host's)
    fnode = (lambda a, b: _f64_boxplus(a, b, jit)) if boxplus else _f64
    # fast_nodes = N > 0: the reference's use_fast_scl=True shortcuts (dec.py:269-306, 354-376) for rate-0 / REP nodes of up
    # to N leaves (the reference itself has no size limit; the CUDA kernel prunes up to 32)
    llr_ch = (np.float32(-1.0) * np.asarray(logits, dtype=np.float32)).astype(np.float64)
    B, n = llr_ch.shape
    L = int(list_size)
    kind = "stable" if stable_sort else None
    state = {
        "u": np.zeros((B, L, n), dtype=np.uint8),
        "pm": np.concatenate([np.zeros((B, 1)), np.full((B, L - 1), LLR_MAX)], axis=1),
        # perm[b, p] = which ORIGINAL row of the live LLR stack path p descends from is handled by
        # physically permuting every live array (simple, obviously correct; this is the checker).
        "stack": [],   # list of [B, L, len] arrays + list of [B, L, len] betas, permuted on forks
    }
    bar = np.arange(B)[:, None]

    def permute_all(par):
        """Every survivor inherits its parent's full state (polar_scl.py:109-120)."""
        state["u"] = state["u"][bar, par]
        for item in state["stack"]:
            for key in list(item.keys()):
                item[key] = item[key][bar, par]

    def rec(a, frame):
        """frame['L']: [B, L, len] LLRs of this node (lives on the stack so forks can permute it).
        Returns beta [B, L, len] for the *current* path order."""
        ln = frame["L"].shape[2]
        if ln == 1:
            x = np.maximum(np.minimum(frame["L"][:, :, 0], LLR_MAX), -LLR_MAX)   # polar_scl.py:81
            if frozen[a]:
                state["pm"] = state["pm"] + _softplus_neg(x, use_log1p, jit)   # u=0, polar_scl.py:82
                return np.zeros((B, L, 1), dtype=np.uint8)
            c0 = state["pm"] + _softplus_neg(x, use_log1p, jit)                 # u_hat = 0
            c1 = state["pm"] + _softplus_neg(-x, use_log1p, jit)                # u_hat = 1
            # reference slot order before the sort: [L paths with u=0 | L paths with u=1]
            cand = np.concatenate([c0, c1], axis=1)
            order = np.argsort(cand, axis=1, kind=kind)[:, :L]                  # polar_scl.py:86-92
            par = order % L
            bit = (order // L).astype(np.uint8)
            state["pm"] = np.take_along_axis(cand, order, axis=1)
            permute_all(par)
            state["u"][:, :, a] = bit
            return bit[:, :, None]
        if fast_nodes and ln <= fast_nodes:
            fr = frozen[a:a + ln]
            xs = np.maximum(np.minimum(frame["L"], LLR_MAX), -LLR_MAX)
            if fr.all():                                                         # dec.py:269-280 (rate-0)
                state["pm"] = state["pm"] + _softplus_neg(xs, use_log1p, jit).sum(axis=-1)
                return np.zeros((B, L, ln), dtype=np.uint8)
            if fr[:-1].all() and not fr[-1]:                                     # dec.py:281-306 (REP)
                c0 = state["pm"] + _softplus_neg(xs, use_log1p, jit).sum(axis=-1)
                c1 = state["pm"] + _softplus_neg(-xs, use_log1p, jit).sum(axis=-1)
                cand = np.concatenate([c0, c1], axis=1)
                order = np.argsort(cand, axis=1, kind=kind)[:, :L]
                par = order % L
                bit = (order // L).astype(np.uint8)
                state["pm"] = np.take_along_axis(cand, order, axis=1)
                permute_all(par)
                state["u"][:, :, a + ln - 1] = bit
                return np.repeat(bit[:, :, None], ln, axis=2)
        h = ln // 2
        Lc = frame["L"]
        left = {"L": fnode(Lc[:, :, :h], Lc[:, :, h:])}                          # polar_scl.py:134-137
        state["stack"].append(left)
        bl = rec(a, left)
        state["stack"].pop()
        keep = {"bl": bl}
        state["stack"].append(keep)
        Lc = frame["L"]                                                          # re-read: may be permuted
        right = {"L": _g64(Lc[:, :, :h], Lc[:, :, h:], keep["bl"].astype(np.float64))}  # :140-144
        state["stack"].append(right)
        br = rec(a + h, right)
        state["stack"].pop()
        state["stack"].pop()
        bl = keep["bl"]
        return np.concatenate([bl ^ br, br], axis=2)                            # polar_scl.py:147-153

    root = {"L": np.broadcast_to(llr_ch[:, None, :], (B, L, n)).copy()}          # polar_scl.py:200
    state["stack"].append(root)
    rec(0, root)
    order = np.argsort(state["pm"], axis=1, kind=kind)                           # polar_scl.py:204
    pm = np.take_along_axis(state["pm"], order, axis=1)
    u = state["u"][bar, order]
    return u, pm


def path_metric(logits, frozen, u):
    """Path metric the reference accumulates along a GIVEN decision vector u [B, n] (polar_scl.py:69-85 applied
    to one path whose leaf decisions are forced): sum over all leaves of log(1 + exp(-(1-2u).clip(llr))), with the
    fp64 f / g tree of polar_scl.py:93-108.  Checker for "is this a legitimate path and is its metric right"."""
    llr = (np.float32(-1.0) * np.asarray(logits, dtype=np.float32)).astype(np.float64)
    u = np.asarray(u, dtype=np.uint8)
    B, n = llr.shape
    pm = np.zeros(B)

    def rec(a, Lc):
        nonlocal pm
        ln = Lc.shape[1]
        if ln == 1:
            x = np.maximum(np.minimum(Lc[:, 0], LLR_MAX), -LLR_MAX)
            bit = u[:, a].astype(np.float64)
            assert not (frozen[a] and u[:, a].any()), "frozen position decided 1"
            pm = pm + np.log(1 + np.exp(-(1 - 2 * bit) * x))
            return u[:, a:a + 1]
        h = ln // 2
        bl = rec(a, _f64(Lc[:, :h], Lc[:, h:]))
        br = rec(a + h, _g64(Lc[:, :h], Lc[:, h:], bl.astype(np.float64)))
        return np.concatenate([bl ^ br, br], axis=1)

    rec(0, llr)
    return pm


def scl_decode(logits, frozen_pos, n, list_size):
    """SCL_Dec.forward (polar_scl.py:210-234): best path = argmin pm, info bits at ascending info_pos."""
    fz = frozen_vec(frozen_pos, n)
    u, pm = scl_decode_full(np.asarray(logits).reshape(-1, n), fz, list_size)
    best = u[:, 0, :]
    return best[:, info_positions(frozen_pos, n)].astype(np.float32)


# ------------------------------------------------------------------------------------------------
# A13  CRC + CRC-aided selection      my_sn/fec/crc.py:38-138, my_sn/fec/polar/dec.py:507-527
# ------------------------------------------------------------------------------------------------
CRC_COEFFS = {  # crc.py:40-45 (3GPP TS 38.212 Sec. 5.1)
    "CRC24A": (24, [24, 23, 18, 17, 14, 11, 10, 7, 6, 5, 4, 3, 1, 0]),
    "CRC24B": (24, [24, 23, 6, 5, 1, 0]),
    "CRC24C": (24, [24, 23, 21, 20, 17, 15, 13, 12, 8, 4, 2, 1, 0]),
    "CRC16": (16, [16, 12, 5, 0]),
    "CRC11": (11, [11, 10, 9, 5, 0]),
    "CRC6": (6, [6, 5, 0]),
}


def crc_remainder(bits, crc_degree):
    """Polynomial long division of bits(x).x^len by the generator (MSB first).  Equals the parity the
    reference's generator-matrix encoder appends (crc.py:54-74,103-105).  bits: [..., k] -> [..., len]."""
    ln, coeffs = CRC_COEFFS[crc_degree]
    poly = np.zeros(ln + 1, dtype=np.uint8)
    for c in coeffs:
        poly[ln - c] = 1                                    # MSB first (crc.py:48-52)
    b = (np.asarray(bits) != 0).astype(np.uint8)
    lead = b.shape[:-1]
    k = b.shape[-1]
    work = np.concatenate([b, np.zeros(lead + (ln,), dtype=np.uint8)], axis=-1)
    for i in range(k):
        sel = work[..., i].astype(bool)
        work[sel, i:i + ln + 1] ^= poly
    return work[..., k:]


def crc_encode(bits, crc_degree):
    """CRCEncoder.forward (crc.py:84-109): append parity."""
    b = (np.asarray(bits) != 0).astype(np.uint8)
    return np.concatenate([b, crc_remainder(b, crc_degree)], axis=-1)


def crc_valid(bits_with_parity, crc_degree):
    """CRCDecoder.forward (crc.py:119-138): re-encodes ALL k bits (payload AND parity) with the
    [k, len] generator matrix and declares valid iff every output parity bit is 0."""
    return crc_remainder(bits_with_parity, crc_degree).sum(axis=-1) == 0


def scl_crc_select(u_list, pm, frozen_pos, n, crc_degree):
    """my_sn/fec/polar/dec.py:507-527: pm += (1-valid).30.k ; argmin ; CRC bits stay in the output.
    u_list [B, L, n] (any path order), pm [B, L] -> (u_info [B,k] fp32, chosen index [B])."""
    info = info_positions(frozen_pos, n)
    k = info.shape[0]
    cand = u_list[:, :, info]
    valid = crc_valid(cand, crc_degree)
    pen = pm + (1.0 - valid.astype(np.float64)) * LLR_MAX * k
    idx = np.argmin(pen, axis=-1)
    return cand[np.arange(cand.shape[0]), idx].astype(np.float32), idx


# ------------------------------------------------------------------------------------------------
# A14  LLR front end        awgn_model.py:33-44, ebno.py:21-23, mapping.py:136-149,225-241, awgn.py:19-29
# ------------------------------------------------------------------------------------------------
def ebnodb2no(ebno_db, n_bits_per_sym, coderate):
    """ebno.py:21-23."""
    ebno = 10.0 ** (ebno_db / 10.0)
    return 1.0 / (ebno * coderate * n_bits_per_sym)


def qpsk_awgn_logits(codewords, noise_re, noise_im, no):
    """Closed form of Mapper -> AWGN -> Demapper for the 2-bit constellation (SURVEY 3.2 [probe]):
    even code bits ride the real axis, odd bits the imaginary axis, bit 0 -> +1/sqrt2, bit 1 -> -1/sqrt2
    (mapping.py:40,139-146); y = x + sqrt(no).noise (awgn.py:27-28, noise already N(0, 1/2) per
    dim, utils.py:12-16); exact log-sum-exp demap over the 2 points per dim (mapping.py:195-205)
    reduces to logit = -2.sqrt2.y/no.  Agreement with the layer stack: <= 6e-6 abs (fp32)."""
    c = np.asarray(codewords, dtype=np.float32)
    B, n = c.shape
    y = np.empty((B, n), dtype=np.float32)
    s = np.float32(np.sqrt(np.float32(no)))
    a = np.float32(1.0 / np.sqrt(2.0))
    y[:, 0::2] = (1 - 2 * c[:, 0::2]) * a + s * np.asarray(noise_re, dtype=np.float32)
    y[:, 1::2] = (1 - 2 * c[:, 1::2]) * a + s * np.asarray(noise_im, dtype=np.float32)
    return (np.float32(-2.0 * np.sqrt(2.0)) * y / np.float32(no)).astype(np.float32)


# ------------------------------------------------------------------------------------------------
# A15  error counting       my_sn/sim.py:7-18
# ------------------------------------------------------------------------------------------------
def count_errors(b, b_hat):
    return int(np.sum(np.asarray(b) != np.asarray(b_hat)))


def count_block_errors(b, b_hat):
    return int(np.sum(np.any(np.asarray(b) != np.asarray(b_hat), axis=-1)))


# ------------------------------------------------------------------------------------------------
# N4  ordered-statistics decoder       my_sn/fec/osd/dec.py:8-192
# ------------------------------------------------------------------------------------------------
def osd_decode(logits, gm, t, dtype=np.float32):
    """Restatement of OSDecoder.forward (dec.py:149-191) for a generator matrix gm [k, n] of 0/1.
    Steps (same order as the reference): clip to +-100 (:155); positions by |llr| descending (:157; equal magnitudes in
    index order -- torch.argsort leaves that open); most-reliable basis by the pivot method (:99-117: row c's pivot =
    its first set column, cleared from every other row); second permutation to [pivots | remaining columns ascending]
    (:118-134); hard decisions llr > 0 on the pivots, re-encoded (:171-174); distance mean_j log(1 + exp(llr_j (1 - 2c_j)))
    (:64-79) in `dtype`; error patterns itertools.combinations(range(k), w) for w = 1 .. t (:57-62), argmin (first) within
    a weight (:93), strictly smaller across weights (:181-184); inverse permutation (:186-188).
    Returns (c_hat uint8 [B, n], distance [B], gap [B]) where gap = (runner-up distance - winning distance) / winning
    distance over ALL tested candidates, computed in float64: a decision whose gap is at rounding level is one the
    reference itself does not reproduce across exp/log/sum implementations."""
    import itertools
    x = np.clip(np.asarray(logits, dtype=np.float32), -100.0, 100.0)
    gm = (np.asarray(gm) != 0).astype(np.uint8)
    k, n = gm.shape
    B = x.shape[0]
    pats = [np.array(list(itertools.combinations(range(k), w)), dtype=np.int64).reshape(-1, w) for w in range(1, t + 1) if w <= k]
    out = np.zeros((B, n), dtype=np.uint8)
    dist = np.zeros(B, dtype=dtype)
    gap = np.zeros(B, dtype=np.float64)
    for b in range(B):
        order = np.argsort(-np.abs(x[b]), kind="stable")
        g = gm[:, order].copy()
        piv = []
        for c in range(k):
            p = int(np.argmax(g[c]))
            piv.append(p)
            hit = g[:, p].astype(bool)
            hit[c] = False
            g[hit] ^= g[c]
        rest = [j for j in range(n) if j not in set(piv)]
        perm2 = np.array(piv + rest, dtype=np.int64)
        g = g[:, perm2]
        order = order[perm2]
        l = x[b, order]

        def distance(cands, dt):
            ll = l.astype(dt)[None, :] * (dt(1) - dt(2) * cands.astype(dt))
            with np.errstate(over="ignore"):
                return np.mean(np.log(dt(1) + np.exp(ll)), axis=1, dtype=dt)

        u = (l[:k] > 0).astype(np.uint8)
        c0 = (u.astype(np.int64) @ g.astype(np.int64) % 2).astype(np.uint8)
        best_c, best_d = c0, distance(c0[None], dtype)[0]
        all_d = [distance(c0[None], np.float64)]
        for ep in pats:
            e = np.zeros((ep.shape[0], n), dtype=np.uint8)
            for col in range(ep.shape[1]):
                e ^= g[ep[:, col]]
            cand = e ^ c0[None, :]
            d = distance(cand, dtype)
            i = int(np.argmin(d))
            if d[i] < best_d:
                best_d, best_c = d[i], cand[i]
            all_d.append(distance(cand, np.float64))
        all_d = np.sort(np.concatenate(all_d))
        gap[b] = (all_d[1] - all_d[0]) / max(all_d[0], 1e-300) if all_d.size > 1 and np.isfinite(all_d[1]) else np.inf
        inv = np.argsort(order)
        out[b] = best_c[inv]
        dist[b] = best_d
    return out, dist, gap
