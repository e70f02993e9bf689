"""The drop-in boundary taken literally: the REFERENCE'S OWN sources -- homo/fhe_image.h, fhe_resize.h,
fhe_decode.h and homo/server_{jpeg,resize,decode}.cpp, compiled unchanged from /root/reference against
include/seal/seal.h (recipe: oracle/Makefile target `ref`, outputs oracle/_ref/; one token of
fhe_decode.h is patched at build time, see the Makefile) -- run their ciphertext arithmetic on the
GPU through libfhe_b200, and every ciphertext they produce equals the oracle's restatement bit for bit.

CPU: the recipe builds (when /root/reference is present) and the programs fail loudly without a
device.  GPU: the binaries built here travel with the snapshot; nothing reads /root/reference at run
time.  These are also the graph-level parity tests at the parameters of BASELINE configs 3 and 5:
n = 8192, k = 4, FractionalEncoder(100, 100)."""
import math
import os
import subprocess

import numpy as np
import pytest

from conftest import PKG_NAME, ROOT
from test_host_resize_decode import (approximated_step_relin_oracle, cubic_relin_oracle, linear_relin_oracle,
                                      resize_bicubic_oracle, sincos_relin_oracle)
from test_server_jpeg import wire_load_all, wire_save

REFDIR = os.path.join(ROOT, "oracle", "_ref")
PROGRAMS = ("ref_graphs", "server_jpeg", "server_resize", "server_decode")
YQT = [16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55, 14, 13, 16, 24, 40, 57, 69, 56, 14, 17, 22, 29, 51, 87,
       80, 62, 18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92, 49, 64, 78, 87, 103, 121, 120, 101, 72, 92,
       95, 98, 112, 100, 103, 99]
UVQT = [17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99, 99, 99, 47, 66] + [99] * 38


def ref_exe(pkg, name):
    pkg.build()
    if os.path.isdir("/root/reference/homo"):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"])
    exe = os.path.join(REFDIR, name)
    if not os.path.exists(exe):
        pytest.skip(f"oracle/_ref/{name} not built (needs /root/reference at build time)")
    return exe


def transparent(o, value):
    """seal.h Encryptor: (Delta * m, 0)"""
    ct = np.zeros(o.ct_words(2), dtype=np.uint64)
    if value != 0:
        o.add_plain(ct, o.encode(value))
    return ct


def test_reference_sources_compile_against_the_shim_and_have_no_cpu_fallback(pkg, tmp_path):
    for name in PROGRAMS:
        ref_exe(pkg, name)
    import torch
    if torch.cuda.is_available():
        pytest.skip("device present: covered by the gpu tests")
    os.makedirs(tmp_path / "keys")
    (tmp_path / "keys" / "params.txt").write_text("8 8 3 16384\n")
    r = subprocess.run([os.path.join(REFDIR, "ref_graphs"), "jpeg", "2048", "64", "32", "/dev/null", "/dev/null"],
                       capture_output=True, text=True)
    assert r.returncode != 0 and "no usable CUDA device" in r.stderr
    # the reference's main() does not catch: the exception of SEALContext's constructor terminates it
    r = subprocess.run([os.path.join(REFDIR, "server_jpeg"), "--cmod", "2048"], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode != 0 and "no usable CUDA device" in r.stderr


@pytest.mark.gpu
def test_reference_fhe_image_h_on_gpu(pkg, oracle_mod, tmp_path):
    """rgb_to_ycc_fhe, encrypted_dct, quantize_fhe of the reference's fhe_image.h, op by op on the GPU"""
    exe = ref_exe(pkg, "ref_graphs")
    n = 4096
    o = oracle_mod.Oracle(n)
    W = o.ct_words(2)
    data = o.fill_uniform(0, 192, 2, 0x1A6E)
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    data.tofile(fin)
    r = subprocess.run([exe, "jpeg", str(n), "100", "100", fin, fout], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and r.stdout.startswith("PASS"), r.stdout + r.stderr
    got = np.fromfile(fout, dtype=np.uint64)
    want = data.copy()
    R, G, B = want[:64 * W], want[64 * W:128 * W], want[128 * W:]
    for i in range(64):
        o.rgb_to_ycc_fhe(R[i * W:(i + 1) * W], G[i * W:(i + 1) * W], B[i * W:(i + 1) * W])
    for blk, quant in ((R, YQT), (G, UVQT), (B, UVQT)):
        o.encrypted_dct(blk)
        o.quantize_fhe(blk, quant)
    assert got.size == want.size and (got == want).all()


@pytest.mark.gpu
def test_reference_fhe_resize_h_on_gpu_at_config3_parameters(pkg, oracle_mod, tmp_path):
    """Cubic (two levels: sizes 4 and 6), Linear and a bicubic ResizeImage of the reference's
    fhe_resize.h at n = 8192, k = 4, encoder (100, 100)"""
    exe = ref_exe(pkg, "ref_graphs")
    n, ow, oh, dw, dh = 8192, 5, 4, 2, 2
    o = oracle_mod.Oracle(n)
    k, W = o.k, o.ct_words(2)
    assert k == 4
    data = o.fill_uniform(0, 5, 2, 0xC0B1C)
    src = o.fill_uniform(100, ow * oh * 3, 2, 0xAB5)
    sk, _ = o.keygen(5)
    dbc = 30
    evk = o.evk_gen(sk, dbc, 1, 41)
    fin, fout, fwire, fkey = (str(tmp_path / f) for f in ("in.bin", "out.bin", "stream.wire", "evk.bin"))
    data.tofile(fin)
    evk.tofile(fkey)
    with open(fwire, "wb") as f:
        for i in range(ow * oh * 3):
            wire_save(f, src[i * W:(i + 1) * W], 2, k, n)
    r = subprocess.run([exe, "resize", str(n), "100", "100", fin, fout, fwire, str(ow), str(oh), str(dw), str(dh), fkey, str(dbc)],
                       capture_output=True, text=True, timeout=1800)
    assert r.returncode == 0 and r.stdout.startswith("PASS"), r.stdout + r.stderr
    got = np.fromfile(fout, dtype=np.uint64)

    def ct(i):
        return data[i * W:(i + 1) * W].copy()

    c1, rs = o.cubic(ct(0), ct(1), ct(2), ct(3), [2, 2, 2, 2], ct(4))
    assert rs == 4
    c2, rs = o.cubic(c1, c1, c1, c1, [4, 4, 4, 4], ct(4))
    assert rs == 6
    lin, rs = o.linear(ct(0), 2, ct(1), 2, ct(4))
    assert rs == 3
    fr = []
    for y in range(dh):
        v = np.float32(np.float64(np.float32(y) / np.float32(dh - 1) * np.float32(oh)) - 0.5)
        for x in range(dw):
            u = np.float32(np.float64(np.float32(x) / np.float32(dw - 1) * np.float32(ow)) - 0.5)
            fr += [transparent(o, float(u - np.floor(u))), transparent(o, float(v - np.floor(v)))]
    res, nf = resize_bicubic_oracle(o, src, ow, oh, dw, dh, np.concatenate(fr))
    assert nf == 2 * dw * dh
    # the same two graphs with every product relinearised (Evaluator::set_relinearization under the reference's code)
    cub_rl = cubic_relin_oracle(o, ct(0), ct(1), ct(2), ct(3), ct(4), evk, dbc)
    lin_rl = linear_relin_oracle(o, ct(0), ct(1), ct(4), evk, dbc)
    want = np.concatenate([c1, c2, lin, res, cub_rl, lin_rl])
    assert got.size == want.size
    assert (got == want).all()


@pytest.mark.gpu
def test_reference_fhe_decode_h_on_gpu_at_config5_parameters(pkg, oracle_mod, tmp_path):
    """homomorphic_sin / cos (size 11) and approximated_step (sizes 21, 22: products 9x2, 11x11, 21x2) of the
    reference's fhe_decode.h at n = 8192, k = 4, encoder (100, 100) -- as written, and with every
    product relinearised (dbc 30: BASELINE config 5)"""
    exe = ref_exe(pkg, "ref_graphs")
    n, order, degree, delta, npos, dbc = 8192, 8, 2, 0.25, 2, 30
    o = oracle_mod.Oracle(n)
    W = o.ct_words(2)
    data = o.fill_uniform(0, 5, 2, 0xDEC0DE)
    sk, _ = o.keygen(3)
    evk = o.evk_gen(sk, dbc, 1, 99)
    fin, fout, fkey = str(tmp_path / "in.bin"), str(tmp_path / "out.bin"), str(tmp_path / "evk.bin")
    data.tofile(fin)
    evk.tofile(fkey)
    r = subprocess.run([exe, "decode", str(n), "100", "100", fin, fout, str(order), str(degree), str(delta), str(npos), fkey,
                        str(dbc)], capture_output=True, text=True, timeout=1800)
    assert r.returncode == 0 and r.stdout.startswith("PASS"), r.stdout + r.stderr
    got = np.fromfile(fout, dtype=np.uint64)

    def ct(i):
        return data[i * W:(i + 1) * W].copy()

    zero = np.zeros(W, dtype=np.uint64)                      # transparent encrypt(encode(0.0))
    zeros = np.tile(zero, 2 * degree * npos)
    s, rs = o.homomorphic_sincos(ct(0), 2, zero, 0)
    c, rc = o.homomorphic_sincos(ct(1), 2, zero, 1)
    assert rs == 11 and rc == 11
    run, rr = o.approximated_step(ct(2), ct(3), ct(4), order, degree, delta, npos, zeros)
    assert rr == 22
    want = [s, c, run]
    want.append(sincos_relin_oracle(o, ct(0), zero, evk, dbc, 1.0))
    want.append(sincos_relin_oracle(o, ct(1), zero, evk, dbc, -1.0))
    want.append(approximated_step_relin_oracle(o, ct(2), ct(3), ct(4), order, degree, delta, npos, zeros, evk, dbc))
    want = np.concatenate(want)
    assert got.size == want.size
    assert (got == want).all()


def _workdir(tmp_path, params):
    os.makedirs(tmp_path / "keys")
    os.makedirs(tmp_path / "image")
    (tmp_path / "keys" / "params.txt").write_text(params)
    return tmp_path


@pytest.mark.gpu
def test_reference_server_jpeg_cpp_on_gpu(pkg, oracle_mod, tmp_path):
    """homo/server_jpeg.cpp, unchanged: two RGB tiles in, Y/Cb/Cr DCT coefficients out"""
    exe = ref_exe(pkg, "server_jpeg")
    n = 2048
    o = oracle_mod.Oracle(n)
    k, W = o.k, o.ct_words(2)
    wd = _workdir(tmp_path, "16 8 3 16384\n")
    data = o.fill_uniform(0, 2 * 192, 2, 0x5E12)
    with open(wd / "image" / "nothingpersonnel.txt", "wb") as f:
        for i in range(2 * 192):
            wire_save(f, data[i * W:(i + 1) * W], 2, k, n)
    r = subprocess.run([exe, "--cmod", str(n)], capture_output=True, text=True, cwd=wd, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("RGBYCC,") == 2 and r.stdout.count("DCT,") == 2
    got = wire_load_all(wd / "image" / "zoop.txt", k, n)
    assert len(got) == 2 * 192
    want = data.copy()
    for t in range(2):
        tile = want[t * 192 * W:(t + 1) * 192 * W]
        R, G, B = tile[:64 * W], tile[64 * W:128 * W], tile[128 * W:]
        for i in range(64):
            o.rgb_to_ycc_fhe(R[i * W:(i + 1) * W], G[i * W:(i + 1) * W], B[i * W:(i + 1) * W])
        for blk in (R, G, B):
            o.encrypted_dct(blk)
    assert (np.concatenate(got) == want).all()


@pytest.mark.gpu
def test_reference_server_resize_cpp_on_gpu(pkg, oracle_mod, tmp_path):
    """homo/server_resize.cpp --bicubic, unchanged, at n = 8192"""
    exe = ref_exe(pkg, "server_resize")
    n, ow, oh, dw, dh = 8192, 4, 4, 2, 2
    o = oracle_mod.Oracle(n)
    k, W = o.k, o.ct_words(2)
    wd = _workdir(tmp_path, f"{ow} {oh} 3 16384\n")
    src = o.fill_uniform(0, ow * oh * 3, 2, 0x5E13)
    with open(wd / "image" / "nothingpersonnel.txt", "wb") as f:
        for i in range(ow * oh * 3):
            wire_save(f, src[i * W:(i + 1) * W], 2, k, n)
    r = subprocess.run([exe, "--bicubic", "-w", str(dw), "-h", str(dh), "--cmod", str(n)], capture_output=True, text=True,
                       cwd=wd, timeout=1800)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.startswith("Cubic,") and len(r.stdout.strip().rstrip(",").split(",")) == 1 + 15 * dw * dh
    got = wire_load_all(wd / "image" / "zoop.txt", k, n)
    fr = []
    for y in range(dh):
        v = np.float32(np.float64(np.float32(y) / np.float32(dh - 1) * np.float32(oh)) - 0.5)
        for x in range(dw):
            u = np.float32(np.float64(np.float32(x) / np.float32(dw - 1) * np.float32(ow)) - 0.5)
            fr += [transparent(o, float(u - np.floor(u))), transparent(o, float(v - np.floor(v)))]
    want, _ = resize_bicubic_oracle(o, src, ow, oh, dw, dh, np.concatenate(fr))
    assert len(got) == dw * dh * 3 and all(g.size == o.ct_words(6) for g in got)
    assert (np.concatenate(got) == want).all()


@pytest.mark.gpu
def test_reference_server_decode_cpp_refuses_the_secret_key_overload(pkg, oracle_mod, tmp_path):
    """homo/server_decode.cpp:131 calls the DECRYPTING overload of approximated_step (it holds the secret
    key "for debugging"); the server library has no secret-key operations, and says so."""
    exe = ref_exe(pkg, "server_decode")
    n = 2048
    o = oracle_mod.Oracle(n)
    k, W = o.k, o.ct_words(2)
    wd = _workdir(tmp_path, "2 1 1 0 0\n")
    recs = o.fill_uniform(0, 2, 2, 0x5E14)
    with open(wd / "image" / "nothingpersonnel.txt", "wb") as f:
        for i in range(2):
            wire_save(f, recs[i * W:(i + 1) * W], 2, k, n)
    r = subprocess.run([exe, "--cmod", str(n)], capture_output=True, text=True, cwd=wd, timeout=900)
    assert r.returncode != 0 and "secret-key operations" in r.stderr
