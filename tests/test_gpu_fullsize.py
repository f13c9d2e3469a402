"""Parity at BASELINE.json's FULL batch sizes, inside the driver-run GPU suite (VERDICT r01 "weak" #1):

  configs[1]  SC  k=512  n=1024  B=2^20            -- every decision of every codeword against the C restatement
  configs[2]  SCL L=8 (501 + CRC11) n=1024 B=2^18  -- best path, best metric AND the CRC-aided selection of all 2^18
  configs[3]  SCL L=32 k=1024 n=2048, 4096 codewords of one sweep iteration

Rule for the list decoder (exact metric ties between distinct paths exist because of the +-30 clip, and the
reference's literal log(1+exp(.)) is evaluated by whatever exp/log the host has): a codeword on which the GPU and the C
restatement disagree must be one the REFERENCE ITSELF does not reproduce -- some variant of the numpy restatement
(log1p instead of log(1+.), unstable instead of stable sort, penalties moved by one ulp at random) must change the
outcome too.  A mismatch on a codeword where all variants agree with each other is a kernel bug and fails the test.
On every mismatch the GPU's answer must still be a legitimate path: its reported metric equals the metric the
reference accumulates along that decision vector (oracle.path_metric)."""
import os

import numpy as np
import pytest

from util import golden, unpack_words, pack_words

pytestmark = pytest.mark.gpu
PM_RTOL = 1e-5
VARIANTS = [dict(use_log1p=True), dict(stable_sort=False), dict(use_log1p=True, stable_sort=False)] + \
           [dict(ulp_jitter_seed=s) for s in range(1, 7)]


def _env():
    import torch
    import d_kernels as dk
    from oracle import polar_oracle as po, c_oracle as co
    return torch, dk, po, co, torch.device("cuda", 0)


def _outcomes(po, logits, fz, L, select):
    """Per codeword: the set of outcomes (bytes of the chosen decision vector) over the oracle variants."""
    outs = [set() for _ in range(logits.shape[0])]
    for kw in [dict()] + VARIANTS:
        u, pm = po.scl_decode_full(logits, fz, L, **kw)
        ch = select(u, pm)
        for i in range(logits.shape[0]):
            outs[i].add(ch[i].tobytes())
    return outs


def _check_mismatches(po, name, idx, logits, fz, L, got_u, got_pm, ref_u, select, cap):
    """idx: codewords where GPU != C restatement.  Each must be non-robust in the reference itself and legitimate."""
    assert len(idx) <= cap, "%s: %d mismatching codewords (cap %d)" % (name, len(idx), cap)
    if not len(idx):
        return
    outs = _outcomes(po, logits[idx], fz, L, select)
    pm_path = po.path_metric(logits[idx], fz, got_u[idx])
    for j, b in enumerate(idx):
        outs[j].add(ref_u[b].tobytes())
        assert len(outs[j]) > 1, "%s: codeword %d differs from the reference although every oracle variant agrees on it" % (name, b)
        if got_pm is not None:        # metric of the GPU's path = what the reference accumulates along it (+30 per dummy ancestry)
            d = got_pm[b] - pm_path[j]
            d -= 30.0 * np.round(d / 30.0)
            assert abs(d) <= PM_RTOL * max(abs(pm_path[j]), 1.0), (name, b, got_pm[b], pm_path[j])


def test_sc_config1_full_batch_bit_exact():
    """configs[1]: all 2^20 codewords (k=512, n=1024, 4 dB) bit-exact against the C restatement (polar_sc.py:54-133)."""
    torch, dk, po, co, dev = _env()
    n, k, B = 1024, 512, 1 << 20
    fp = golden("frozen_sets")["rm_1024_512"]
    tables = dk.code_tables(fp, n, dev)
    _, _, x = dk.awgn_frontend(tables, B, po.ebnodb2no(4.0, 2, k / n), 1234)
    u_info, u_packed = dk.sc_decode(x, tables, want_info=False, want_packed=True)
    fz = po.frozen_vec(fp, n)
    bad = 0
    step = 1 << 18
    for a in range(0, B, step):
        ref = co.sc_decode_full(x[a:a + step].cpu().numpy(), fz)
        bad += int((pack_words(ref) != u_packed[a:a + step].cpu().numpy().view(np.uint32)).any(axis=1).sum())
    assert bad == 0, "%d of %d codewords differ" % (bad, B)


def test_scl_config2_full_batch_crc_aided():
    """configs[2]: SCL L=8, k=512 = 501 payload + CRC11, n=1024, all 2^18 codewords at 3 dB: best path bit-exact, best
    metric within 1e-5, CRC-aided selection (dec.py:507-527) identical -- up to codewords the reference does not
    reproduce itself (module docstring)."""
    torch, dk, po, co, dev = _env()
    from my_sn.fec.crc import CRCEncoder
    n, k, L, B = 1024, 512, 8, 1 << 18
    deg = "CRC11"
    fp = golden("frozen_sets")["rm_1024_512"]
    fz = po.frozen_vec(fp, n)
    info = po.info_positions(fp, n)
    tables = dk.code_tables(fp, n, dev)
    chk = CRCEncoder(deg, k)
    gen = CRCEncoder(deg, k - chk.crc_length)
    rows = torch.from_numpy(chk.syndrome_rows(tables.info_pos_np, n).view(np.int32).copy()).to(dev)
    g = torch.Generator(device=dev); g.manual_seed(20)
    payload = torch.randint(0, 2, (B, k - gen.crc_length), device=dev, dtype=torch.float32, generator=g)
    lg = dk.qpsk_awgn_llr(dk.encode_f32(gen(payload), tables), po.ebnodb2no(3.0, 2, k / n), 4321)
    plain = dk.scl_decode(lg, tables, L, want_info=False, want_packed=True, want_pm=True)
    aided = dk.scl_decode(lg, tables, L, crc_rows=rows, crc_len=chk.crc_length, want_info=False, want_packed=True)["u_packed"]
    best_gpu = plain["u_packed"].cpu().numpy().view(np.uint32)
    pm_gpu = plain["pm"].cpu().numpy()
    aided_gpu = aided.cpu().numpy().view(np.uint32)
    # CRC validity of a candidate = remainder of ALL k decoder outputs is zero (crc.py:119-138); linear -> one matmul
    gmat = po.crc_remainder(np.eye(k, dtype=np.uint8), deg).astype(np.int32)                     # [k, 11]
    x_host = lg.cpu().numpy()
    bad_best, bad_sel, rel_max = [], [], 0.0
    ref_best = np.zeros((B, n // 32), dtype=np.uint32)
    ref_sel = np.zeros((B, n // 32), dtype=np.uint32)
    step = 1 << 14
    for a in range(0, B, step):
        u_ref, pm_ref = co.scl_decode_full(x_host[a:a + step], fz, L)
        ref_best[a:a + step] = pack_words(u_ref[:, 0])
        rel = np.abs(pm_gpu[a:a + step, 0] - pm_ref[:, 0]) / np.maximum(np.abs(pm_ref[:, 0]), 1e-30)
        same = (best_gpu[a:a + step] == ref_best[a:a + step]).all(axis=1)
        rel_max = max(rel_max, float(rel[same].max()))
        cand = u_ref[:, :, info]
        valid = ((cand.reshape(-1, k).astype(np.int32) @ gmat) & 1).sum(axis=1).reshape(-1, L) == 0
        idx = np.argmin(pm_ref + (1.0 - valid) * 30.0 * k, axis=1)                                # dec.py:517-520
        ref_sel[a:a + step] = pack_words(u_ref[np.arange(u_ref.shape[0]), idx])
    bad_best = np.nonzero((best_gpu != ref_best).any(axis=1))[0]
    bad_sel = np.nonzero((aided_gpu != ref_sel).any(axis=1))[0]
    assert rel_max <= PM_RTOL, rel_max
    cap = max(4, B // 20000)

    def sel_best(u, pm):
        return u[:, 0]

    def sel_crc(u, pm):
        valid = po.crc_valid(u[:, :, info], deg)
        idx = np.argmin(pm + (1.0 - valid) * 30.0 * k, axis=1)
        return u[np.arange(u.shape[0]), idx]
    _check_mismatches(po, "configs[2] best path", bad_best, x_host, fz, L, unpack_words(best_gpu, n), pm_gpu[:, 0],
                      unpack_words(ref_best, n), sel_best, cap)
    _check_mismatches(po, "configs[2] CRC-aided selection", bad_sel, x_host, fz, L, unpack_words(aided_gpu, n), None,
                      unpack_words(ref_sel, n), sel_crc, cap)
    print("configs[2]: %d codewords, best path differs on %d, CRC-aided selection on %d (all non-robust in the reference), "
          "max rel err of the best metric %.2e" % (B, len(bad_best), len(bad_sel), rel_max))


def test_scl_config3_sample():
    """configs[3]: SCL L=32, k=1024, n=2048 -- 4096 codewords of one sweep iteration (3.5 dB) against the C restatement."""
    torch, dk, po, co, dev = _env()
    n, k, L, B = 2048, 1024, 32, 1 << 12
    fp = golden("frozen_sets")["rm_2048_1024"]
    fz = po.frozen_vec(fp, n)
    tables = dk.code_tables(fp, n, dev)
    _, _, lg = dk.awgn_frontend(tables, B, po.ebnodb2no(3.5, 2, k / n), seed=33)
    res = dk.scl_decode(lg, tables, L, want_info=False, want_packed=True, want_pm=True)
    x_host = lg.cpu().numpy()
    u_ref, pm_ref = co.scl_decode_full(x_host, fz, L)
    got = res["u_packed"].cpu().numpy().view(np.uint32)
    pm = res["pm"].cpu().numpy()
    ref_best = pack_words(u_ref[:, 0])
    same = (got == ref_best).all(axis=1)
    rel = np.abs(pm[:, 0] - pm_ref[:, 0]) / np.maximum(np.abs(pm_ref[:, 0]), 1e-30)
    assert rel[same].max() <= PM_RTOL
    _check_mismatches(po, "configs[3] best path", np.nonzero(~same)[0], x_host, fz, L, unpack_words(got, n), pm[:, 0],
                      u_ref[:, 0], lambda u, p: u[:, 0], cap=2)


def test_scl_known_ill_conditioned_codeword():
    """Round-1 caveat: codeword 122257 of the n=1024 L=8 B=2^17 batch (seed 4242, 3 dB) is the one best path in
    131072 that differed from the CPU restatements.  It is regenerated here (the front end is a pure function of seed and
    codeword index), and must be (a) a legitimate path whose metric the oracle reproduces, (b) non-robust in the oracle."""
    torch, dk, po, co, dev = _env()
    n, k, L = 1024, 512, 8
    fp = po.rm_frozen_pos(n, n - k)
    fz = po.frozen_vec(fp, n)
    tables = dk.code_tables(fp, n, dev)
    _, _, x = dk.awgn_frontend(tables, 8, po.ebnodb2no(3.0, 2, k / n), 4242, offset=122257 - 3)
    res = dk.scl_decode(x, tables, L, want_info=False, want_packed=True, want_pm=True)
    xh = x.cpu().numpy()
    u_ref, pm_ref = co.scl_decode_full(xh, fz, L)
    got = unpack_words(res["u_packed"].cpu().numpy(), n)
    pm = res["pm"].cpu().numpy()
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out_dir):      # kept as a CPU fixture (tests/golden/illcond_scl_1024_L8.npz, tests/test_oracle_golden.py)
        np.savez(os.path.join(out_dir, "cw122257.npz"), logits=xh, gpu_best=got, gpu_pm=pm, c_best=u_ref[:, 0], c_pm=pm_ref)
    bad = np.nonzero((got != u_ref[:, 0]).any(axis=1))[0]
    _check_mismatches(po, "codeword 122257", bad, xh, fz, L, got, pm[:, 0], u_ref[:, 0], lambda u, p: u[:, 0], cap=1)
    ok = np.setdiff1d(np.arange(8), bad)
    assert np.allclose(pm[ok, 0], pm_ref[ok, 0], rtol=PM_RTOL)
