"""CPU tests of the host layer and of the C-ABI library surface (no compute without a GPU)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import oracle as O
import workloads as W
import sahara_b200 as sb
from sahara_b200 import _native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared(header, prefix):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(" + prefix + r"_[a-z0-9_]+)\s*\(", text)))


def exported(lib):
    out = subprocess.check_output(["nm", "-D", "--defined-only", os.path.join(ROOT, "sahara_b200", lib)], text=True)
    return set(line.split()[-1] for line in out.splitlines() if " T " in line)


def test_cuda_library_exports_every_declared_symbol():
    names = declared("sahara_b200.h", "sb200")
    assert len(names) >= 30
    exp = exported("libsahara_b200.so")
    assert [n for n in names if n not in exp] == []
    assert sorted(N.SB200_SYMBOLS) == names  # the ctypes table binds exactly the header
    assert N.cuda.sb200_abi_version() == 3


def test_host_library_exports_every_declared_symbol():
    names = declared("sahara_host.h", "sbh")
    exp = exported("libsahara_host.so")
    assert [n for n in names if n not in exp] == []
    assert sorted(N.SBH_SYMBOLS) == names


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(sb.SaharaError, match="no CPU fallback|CUDA"):
        sb.Context(0)


def test_product_does_not_touch_the_oracle():
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "sahara_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                text = open(os.path.join(base, f), errors="replace").read()
                if re.search(r"(import oracle|from oracle|libsahara_oracle|orc_[a-z_]+\()", text):
                    bad.append(f)
    assert bad == []


def test_fasta_alphabet_and_revcomp(tmp_path):
    p = os.path.join(tmp_path, "q.fa")
    with open(p, "w") as f:
        f.write(">r1 first\nACGT\nacgtn\n>r2\nTTTTU\r\n\n>r3\nN\n")
    r = sb.load_fasta_ranks(p, sigma=6)
    assert [x.tolist() for x in r] == [[1, 2, 3, 4, 1, 2, 3, 4, 5], [4, 4, 4, 4, 4], [5]]
    both = sb.load_fasta_ranks(p, sigma=6, with_revcomp=True)
    assert both[1].tolist() == [5, 1, 2, 3, 4, 1, 2, 3, 4] and both[3].tolist() == [1, 1, 1, 1, 1]
    assert sb.revcomp_ranks(np.array([1, 2, 5, 0, 4], np.uint8)).tolist() == [1, 0, 5, 3, 4]
    with pytest.raises(sb.SaharaError, match="invalid character at position 8"):
        sb.load_fasta_ranks(p, sigma=5)  # N is not in d_dna4
    with open(p, "w") as f:
        f.write(">x\nACGTXA\n")
    with pytest.raises(sb.SaharaError, match=r"query 'x' \(1\) has invalid character at position 4 'X'\(58\)"):
        sb.load_fasta_ranks(p, sigma=6)


def test_parallel_read_set_loader_matches_sequential_reader(tmp_path):
    """`sahara search` loads its reads with a reader that cuts the file into one piece per thread
    (host/fasta.hpp readRanksParallel): same ranks as the sequential reader for every thread count and line layout,
    and the first problem in file order is the one reported."""
    rng = np.random.default_rng(11)
    n, m = 257, 37
    reads = rng.integers(1, 6, size=(n, m))
    p = os.path.join(tmp_path, "reads.fa")
    with open(p, "w", newline="") as f:
        f.write("\n\n")
        for i in range(n):
            txt = "".join("$ACGTN"[c] for c in reads[i])
            eol = "\r\n" if i % 5 == 0 else "\n"
            if i % 3 == 0:  # sequence over several lines, a blank line in between
                f.write(f">r{i} x{eol}{txt[:10]}{eol}{eol}{txt[10:]}{eol}")
            else:
                f.write(f">r{i}{eol}{txt}{eol}")
    want = np.stack(sb.load_fasta_ranks(p, sigma=6))
    for threads in (1, 2, 3, 7, 16, 64):
        got = sb.load_fasta_reads(p, sigma=6, threads=threads)
        assert got.shape == (n, m) and np.array_equal(got, want)
    lines = open(p, newline="").read()
    # an invalid character in record 200, a short record 100 and a short record 230: record 100 is reported
    bad = lines.replace(">r200\r\n", ">r200\r\nX", 1)
    i100 = bad.index(">r100")
    j100 = bad.index("\n", bad.index("\n", i100) + 1)
    bad = bad[:j100 - 2] + bad[j100 - 1:]  # drop the last base (the line ends with \r\n)
    open(p, "w", newline="").write(bad)
    for threads in (1, 4, 16):
        with pytest.raises(sb.SaharaError, match=r"query 'r100' has length 36, .* first query \(37\)"):
            sb.load_fasta_reads(p, sigma=6, threads=threads)
    # only the invalid character: reference message with the 1-based record number
    open(p, "w", newline="").write(lines.replace(">r200\r\n", ">r200\r\nX", 1))
    for threads in (1, 4, 16):
        with pytest.raises(sb.SaharaError, match=r"query 'r200' \(201\) has invalid character at position 0 'X'\(58\)"):
            sb.load_fasta_reads(p, sigma=6, threads=threads)
    open(p, "w").write("ACGT\n>r\nACGT\n")
    with pytest.raises(sb.SaharaError, match="does not start with a '>' header"):
        sb.load_fasta_reads(p, sigma=6, threads=3)
    open(p, "w").write("")
    assert sb.load_fasta_reads(p, sigma=6, threads=3).shape[0] == 0


def test_index_file_reader_agrees_with_oracle_and_rejects_garbage(tmp_path):
    rng = np.random.default_rng(3)
    seqs = [W.random_genome(rng, 5000, with_n=True), W.random_genome(rng, 100)]
    ix = O.OracleIndex.build(seqs, 6, 16)
    path = os.path.join(tmp_path, "a.idx")
    ix.save(path)
    sigma = C.c_uint64()
    N.check_host(N.host.sbh_idx_peek_sigma(path.encode(), C.byref(sigma)))
    assert sigma.value == 6
    view, handle = N.IndexView(), C.c_void_p()
    N.check_host(N.host.sbh_idx_load(path.encode(), C.byref(view), C.byref(handle)))
    info = ix.info()
    assert (view.sigma, view.n_rows, view.n_ssa, view.sampling_rate, view.bits_for_position) == (
        6, info["n_rows"], info["n_ssa"], 16, info["bits_for_position"])
    out = os.path.join(tmp_path, "b.idx")
    N.check_host(N.host.sbh_idx_save(out.encode(), C.byref(view)))
    N.host.sbh_idx_free(handle)
    data = open(path, "rb").read()
    assert open(out, "rb").read() == data
    # the oracle ingests the same in-memory image
    view2, handle2 = N.IndexView(), C.c_void_p()
    N.check_host(N.host.sbh_idx_load(path.encode(), C.byref(view2), C.byref(handle2)))
    again = O.OracleIndex.from_view(view2)
    N.host.sbh_idx_free(handle2)
    assert np.array_equal(again.bwt(0), ix.bwt(0)) and np.array_equal(again.bwt(1), ix.bwt(1))
    for bad, msg in [(data[:-9], "unexpected end of file|not understood"), (data + b"x", "trailing bytes"),
                     ((7).to_bytes(8, "little") + data[8:], "unknown index with 7 letters"),
                     (data[:8] + (5).to_bytes(8, "little") + data[16:], "not understood")]:
        p = os.path.join(tmp_path, "bad.idx")
        open(p, "wb").write(bad)
        v, h = N.IndexView(), C.c_void_p()
        assert N.host.sbh_idx_load(p.encode(), C.byref(v), C.byref(h)) != 0
        assert re.search(msg, N.host.sbh_last_error().decode())
        with pytest.raises(O.OracleError):
            O.OracleIndex.load(p)
    assert N.host.sbh_idx_load(b"/nonexistent.idx", C.byref(view), C.byref(handle)) != 0
    assert "no valid index path" in N.host.sbh_last_error().decode()


def test_index_file_reader_accepts_64_byte_aligned_blocks(tmp_path):
    """upstream declares the occurrence block alignas(64); should it be serialized with its padding, the loader finds the
    stride from the position of the superblock count (layout::blockBytesPadded)"""
    rng = np.random.default_rng(4)
    ix = O.OracleIndex.build([W.random_genome(rng, 3000)], 6, 16)
    path = os.path.join(tmp_path, "a.idx")
    ix.save(path)
    data = open(path, "rb").read()
    view, handle = N.IndexView(), C.c_void_p()
    N.check_host(N.host.sbh_idx_load(path.encode(), C.byref(view), C.byref(handle)))
    nb, sigma = view.n_blocks, 6
    N.host.sbh_idx_free(handle)
    packed = 10 * sigma

    def pad_occ(buf, at):  # one occurrence table at offset `at`: u64 count, blocks, ... -> (bytes with 64-byte blocks, end of the blocks)
        assert int.from_bytes(buf[at:at + 8], "little") == nb
        blocks = np.frombuffer(buf[at + 8:at + 8 + nb * packed], np.uint8).reshape(nb, packed)
        wide = np.zeros((nb, 64), np.uint8)
        wide[:, :packed] = blocks
        return buf[at:at + 8] + wide.tobytes(), at + 8 + nb * packed

    first, end1 = pad_occ(data, 8)
    n_super = (nb + 1023) // 1024
    tail1 = 8 + n_super * sigma * 8 + 8   # superblock vector + rows
    second, end2 = pad_occ(data, end1 + tail1)
    padded = data[:8] + first + data[end1:end1 + tail1] + second + data[end2:]
    p2 = os.path.join(tmp_path, "padded.idx")
    open(p2, "wb").write(padded)
    view2, handle2 = N.IndexView(), C.c_void_p()
    N.check_host(N.host.sbh_idx_load(p2.encode(), C.byref(view2), C.byref(handle2)))
    again = O.OracleIndex.from_view(view2)
    N.host.sbh_idx_free(handle2)
    assert np.array_equal(again.bwt(0), ix.bwt(0)) and np.array_equal(again.bwt(1), ix.bwt(1))


def test_record_decoder_reads_fixed_and_delta_coded_results():
    """sbh_decode_records against a Python encoder of both CSR forms of sb200_wait_batch (include/sahara_b200.h)"""
    rng = np.random.default_rng(9)
    bits, rec = 32, 5
    nq = 300
    hits, ends_fixed, fixed, ends_delta, delta = [], [], bytearray(), [], bytearray()
    for q in range(nq):
        n = int(rng.integers(0, 12)) if q % 7 else 0
        base = int(rng.integers(0, 1 << 31))
        vals = sorted(((int(rng.integers(0, 3)) << bits | (base + int(rng.integers(0, 40)) if rng.random() < 0.8 else int(rng.integers(0, 1 << 32)))) << 4)
                      | int(rng.integers(0, 3)) for _ in range(n))
        prev = None
        for v in vals:
            hits.append((q + 10, (v >> 4) >> bits, (v >> 4) & ((1 << bits) - 1), v & 15))
            fixed += v.to_bytes(rec, "little")
            if prev is None:
                delta += v.to_bytes(rec, "little")
            else:
                d = v - prev
                while d >= 128:
                    delta.append((d & 0x7f) | 0x80)
                    d >>= 7
                delta.append(d)
            prev = v
        ends_fixed.append(len(fixed) // rec)
        ends_delta.append(len(delta))
    want = np.array(hits, dtype=np.uint64).reshape(-1, 4)
    for ends, blob, coded in ((ends_fixed, fixed, 0), (ends_delta, delta, 1)):
        e = np.array(ends, dtype=np.uint32)
        b = np.frombuffer(bytes(blob) + b"\0" * 8, dtype=np.uint8)
        out = np.zeros((want.shape[0], 4), dtype=np.uint64)
        N.check_host(N.host.sbh_decode_records(e.ctypes.data, b.ctypes.data, nq, want.shape[0], rec, bits, coded, 10, out.ctypes.data))
        assert np.array_equal(out, want)
        # a wrong hit count is refused
        assert N.host.sbh_decode_records(e.ctypes.data, b.ctypes.data, nq, want.shape[0] + 1, rec, bits, coded, 10,
                                         np.zeros((want.shape[0] + 1, 4), np.uint64).ctypes.data) != 0
    assert len(delta) < len(fixed)


def test_two_bit_packer():
    """sbh_pack_reads2: A, C, G, T at 2 bits per base, 16 per word; any other symbol is refused with the read's index"""
    rng = np.random.default_rng(12)
    for m in (1, 15, 16, 17, 150, 250):
        r = rng.integers(1, 5, size=(37, m), dtype=np.uint8)
        p = sb.pack_reads2(r, threads=3)
        assert p.shape == (37, (m + 15) // 16)
        codes = np.zeros((37, p.shape[1] * 16), dtype=np.uint8)
        for j in range(16):
            codes[:, j::16] = (p >> np.uint32(2 * j)) & np.uint32(3)
        assert np.array_equal(codes[:, :m] + 1, r) and not codes[:, m:].any()
    bad = rng.integers(1, 5, size=(9000, 40), dtype=np.uint8)
    bad[7123, 5] = 5
    with pytest.raises(sb.SaharaError, match="read 7123 holds a symbol other than A, C, G, T"):
        sb.pack_reads2(bad, threads=4)


def test_synth_mirror_is_deterministic():
    from sahara_b200 import synth
    g = synth.genome(5000, 42)
    assert g.min() >= 1 and g.max() <= 4 and abs(np.bincount(g)[1:] / 5000 - 0.25).max() < 0.03
    r = synth.reads(g, 20, 50, 2, True, 43)
    assert r.shape == (40, 50) and np.array_equal(r[1], (5 - r[0])[::-1])
    assert np.array_equal(r, synth.reads(g, 20, 50, 2, True, 43))
    assert np.array_equal(synth.reads(g, 5, 50, 2, True, 43, first_read=10), r[20:30])
