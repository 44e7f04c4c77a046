"""CPU-side tests (no GPU needed): the oracle is pinned against the reference's own golden values and our committed
fixtures; the host simulation of the kernel headers (tests/hostsim, a test tool) is diffed against both; the C ABI
library loads, exports every symbol include/*.h declares and refuses to run without a CUDA device."""
import ctypes as C
import glob
import os
import re
import subprocess
import sys
import zlib

import numpy as np
import pytest

import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz")))
HAVE_REF = os.path.exists(O.REF_SO) or os.path.isdir("/root/reference/opus-fix")
needs_ref = pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not built and reference sources absent")


def _hostsim():
    so = os.path.join(ROOT, "tests", "hostsim", "libhostsim.so")
    src = os.path.join(ROOT, "tests", "hostsim", "hostsim.cpp")
    newest = max(os.path.getmtime(p) for p in glob.glob(os.path.join(ROOT, "concentus_b200", "csrc", "*")) + [src])
    if not os.path.exists(so) or os.path.getmtime(so) < newest:
        subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-Wno-unused-function", "-o", so, src], check=True)
    hs = C.CDLL(so)
    hs.hostsim_decode_stream.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                         C.c_void_p, C.c_void_p]
    return hs


def _hostsim_decode(hs, data, offs, lens, fs, ch):
    F = len(lens)
    pcm = np.zeros((F * fs, ch), dtype=np.int16)
    rng = np.zeros(F, dtype=np.uint32)
    ret = np.zeros(F, dtype=np.int32)
    hs.hostsim_decode_stream(O.ptr(data), O.ptr(np.ascontiguousarray(offs, dtype=np.int64)), O.ptr(np.ascontiguousarray(lens, dtype=np.int32)),
                             F, fs, ch, 48000, O.ptr(pcm), O.ptr(rng), O.ptr(ret))
    return pcm, rng, ret


def _check_against_golden(g, pcm, rng, ret):
    fs = int(g["frame_size"])
    assert np.array_equal(ret, g["rets"])
    assert np.array_equal(rng, g["dec_ranges"])
    crc = np.array([zlib.crc32(pcm[f * fs:(f + 1) * fs].tobytes()) for f in range(len(ret))], dtype=np.uint32)
    assert np.array_equal(crc, g["pcm_crc"])
    assert np.array_equal(pcm[:2 * fs], g["pcm_head"])


# reference KAT: sum of final ranges over all 65,536 3-byte prefixes of a 4-byte packet, CELT configs 16/20/24/28
# (opus-fix/tests/test_opus_decode.c:236-258)
CRES = {16: 116290185, 20: 2172123586, 24: 2172123586, 28: 2172123586}


def cres_packets(cfg):
    i = np.arange(65536, dtype=np.uint32)
    pk = np.stack([np.full(65536, cfg << 3, dtype=np.uint32), i >> 8, i & 255, np.full(65536, 255, dtype=np.uint32)], axis=1)
    data = pk.astype(np.uint8).reshape(-1)
    return data, np.arange(65536, dtype=np.int64) * 4, np.full(65536, 4, dtype=np.int32)


@needs_ref
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_reproduces_golden(path):
    g = np.load(path)
    pcm, rng, ret = O.decode_stream(g["data"], g["offs"], g["lens"], int(g["frame_size"]), int(g["channels"]))
    _check_against_golden(g, pcm, rng, ret)
    # encoder and decoder of the reference agree on the final range of every packet (tests/test_opus_encode.c:306)
    assert np.array_equal(g["enc_ranges"], g["dec_ranges"])


@needs_ref
@pytest.mark.parametrize("cfg", [16, 28])
def test_oracle_pinned_by_reference_cres_kat(cfg):
    data, offs, lens = cres_packets(cfg)
    _, rng, ret = O.decode_stream(data, offs, lens, 120, 1)
    assert (ret == 120).all()
    assert int(rng.astype(np.uint64).sum() & 0xFFFFFFFF) == CRES[cfg]


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_hostsim_matches_golden(path):
    g = np.load(path)
    pcm, rng, ret = _hostsim_decode(_hostsim(), g["data"], g["offs"], g["lens"], int(g["frame_size"]), int(g["channels"]))
    _check_against_golden(g, pcm, rng, ret)


@pytest.mark.parametrize("cfg", [16, 20, 24, 28])
def test_hostsim_cres_kat(cfg):
    data, offs, lens = cres_packets(cfg)
    _, rng, ret = _hostsim_decode(_hostsim(), data, offs, lens, 120, 1)
    assert (ret == 120).all()
    assert int(rng.astype(np.uint64).sum() & 0xFFFFFFFF) == CRES[cfg]


@needs_ref
def test_hostsim_vs_oracle_mini_sweep():
    hs = _hostsim()
    seed = 77
    for kind in ("music", "clicks"):
        for ch in (1, 2):
            for fs in (120, 480, 960):
                for br, vbr, cvbr in ((32000, 1, 0), (96000, 0, 0), (256000, 1, 1)):
                    seed += 1
                    x = O.test_signal(24000, ch, seed, kind)
                    d, o, l, _ = O.encode_stream(x, fs, br, vbr=vbr, cvbr=cvbr)
                    rp, rr, rret = O.decode_stream(d, o, l, fs, ch)
                    hp, hr, hret = _hostsim_decode(hs, d, o, l, fs, ch)
                    assert np.array_equal(rret, hret) and np.array_equal(rr, hr) and np.array_equal(rp, hp), (kind, ch, fs, br)


def test_hostsim_garbage_packets_do_not_crash_and_match_oracle():
    hs = _hostsim()
    rs = np.random.RandomState(5)
    F = 400
    lens = rs.randint(3, 60, size=F).astype(np.int32)   # >= 2 payload bytes: a 1-byte payload is a PLC frame
    offs = np.zeros(F, dtype=np.int64)
    offs[1:] = np.cumsum(lens[:-1])
    data = rs.randint(0, 256, size=int(lens.sum())).astype(np.uint8)
    # CELT-only, code 0, 2.5..20 ms, mono/stereo TOCs
    for f in range(F):
        data[offs[f]] = 0x80 | (rs.randint(0, 16) << 3) | (rs.randint(0, 2) << 2)
    hp, hr, hret = _hostsim_decode(hs, data, offs, lens, 960, 2)
    assert ((hret > 0) | (hret == -3)).all()    # garbage may overrun its bit budget: OPUS_INTERNAL_ERROR like the reference
    if HAVE_REF:
        rp, rr, rret = O.decode_stream(data, offs, lens, 960, 2)
        assert np.array_equal(rret, hret) and np.array_equal(rr, hr)
        for f in range(F):
            n = max(int(hret[f]), 0)
            assert np.array_equal(rp[f * 960:f * 960 + n], hp[f * 960:f * 960 + n]), f


def _declared_symbols():
    names = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        txt = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        for m in re.finditer(r"\b(opus_[a-z0-9_]+)\s*\(", txt):
            names.add(m.group(1))
    return sorted(names)


def test_cabi_library_loads_and_exports_every_declared_symbol():
    import concentus_b200 as cb
    L = cb.lib()
    syms = _declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(L, s), "include/*.h declares %s but the library does not export it" % s
    assert b"-fixed" in L.opus_get_version_string()
    assert L.opus_strerror(-4) == b"corrupted stream"


def test_host_side_packet_helpers_and_state_block():
    import concentus_b200 as cb
    L = cb.lib()
    toc = np.array([0xFC, 1, 2, 3], dtype=np.uint8)
    assert L.opus_packet_get_bandwidth(O.ptr(toc)) == 1105
    assert L.opus_packet_get_nb_channels(O.ptr(toc)) == 2
    assert L.opus_packet_get_samples_per_frame(O.ptr(toc), 48000) == 960
    assert L.opus_packet_get_nb_frames(O.ptr(toc), 4) == 1
    assert L.opus_packet_get_nb_samples(O.ptr(toc), 4, 48000) == 960
    for ch in (1, 2):
        n = L.opus_decoder_get_size(ch)
        assert 2048 < n <= (1 << 16)          # bounds checked by opus-fix/tests/test_opus_api.c:106
    assert L.opus_decoder_get_size(0) == 0 and L.opus_decoder_get_size(3) == 0
    err = C.c_int(0)
    h = L.opus_decoder_create(48000, 2, C.byref(err))
    assert h and err.value == 0
    v = C.c_int32(0)
    assert L.opus_decoder_ctl(C.c_void_p(h), cb.OPUS_GET_SAMPLE_RATE_REQUEST, C.byref(v)) == 0 and v.value == 48000
    assert L.opus_decoder_ctl(C.c_void_p(h), cb.OPUS_SET_GAIN_REQUEST, C.c_int32(40000)) == cb.OPUS_BAD_ARG
    assert L.opus_decoder_ctl(C.c_void_p(h), 12345) == cb.OPUS_UNIMPLEMENTED
    assert L.opus_decoder_ctl(C.c_void_p(h), cb.OPUS_RESET_STATE) == 0
    L.opus_decoder_destroy(C.c_void_p(h))


def test_product_fails_loudly_without_cuda():
    """No CPU fallback: on a box without a GPU every codec call must return OPUS_INTERNAL_ERROR."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    code = ("import ctypes as C, numpy as np, concentus_b200 as cb; L=cb.lib(); e=C.c_int(0);"
            "h=L.opus_decoder_create(48000,2,C.byref(e)); p=np.array([0xFC,1,2,3],dtype=np.uint8); o=np.zeros(1920,dtype=np.int16);"
            "r=L.opus_decode(C.c_void_p(h), p.ctypes.data_as(C.c_void_p), 4, o.ctypes.data_as(C.c_void_p), 960, 0); print(r, L.opus_b200_init(0))")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT)
    assert out.stdout.split() == ["-3", "-3"], (out.stdout, out.stderr)


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from concentus_b200.shard import shard_range
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    b, e = shard_range(4099, world, rank)
    t = torch.tensor([e - b, b, e], dtype=torch.int64)
    out = [torch.zeros(3, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(out, t)
    ms = torch.tensor([10.0 + rank])
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)        # the timing reduction bench.py does
    dist.barrier()
    if rank == 0:
        q.put(([o.tolist() for o in out], float(ms.item())))
    dist.destroy_process_group()


def test_stream_sharding_world_size_2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    ps = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res, ms = q.get(timeout=120)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == 4099      # contiguous, disjoint, complete
    assert res[0][0] + res[1][0] == 4099 and abs(res[0][0] - res[1][0]) <= 1
    assert ms == 11.0
