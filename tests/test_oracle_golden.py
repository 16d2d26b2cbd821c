"""The oracle against every known-answer vector the reference's own tests hold for the
hot path.  CPU only."""
import numpy as np
import pytest

from oracle import oracle as O
from tests import helpers as H


@pytest.mark.parametrize("name", ["data1", "data2", "data3"])
def test_build_counts_pinned_by_reference(golden, tmp_path, name):
    """tests/build.rs:27-28,46-89 -- k=40, reverse_complement=false, threshold 0"""
    ref = golden["reference_pinned"][name]
    fq = H.write_fastq(tmp_path / f"{name}.fastq", H.golden_seqs(name))
    g, nbytes = O.OracleGIR.create(40, [fq], "fastq", False)
    assert nbytes == ref["bytes"]
    assert g.counts() == (ref["nodes"], ref["edges"])
    st = g.collection_stats()
    for key in ("max_edge_weight", "max_in_degree", "max_out_degree", "incoming_vert_count", "outgoing_vert_count"):
        assert st[key] == ref[key], key
    # CollectionStats equality rounds floats to 2 dp (stats/collections.rs:71-89)
    assert round(st["avg_edge_weight"], 2) == ref["avg_edge_weight"]
    assert round(st["avg_out_degree"], 2) == ref["avg_out_degree"]


def test_filter_pinned_by_reference(golden, tmp_path):
    """tests/pruner.rs:240-251, 260-262, 272"""
    for case in golden["reference_pinned_filter"]:
        fq = H.write_fastq(tmp_path / "f.fastq", H.golden_seqs(case["file"]))
        g, _ = O.OracleGIR.create(40, [fq], "fastq", False)
        g.remove_weak_edges(case["threshold"])
        assert g.counts() == (case["nodes"], case["edges"]), case


def test_too_short_read_aborts(golden, tmp_path):
    """hm_gir.rs:40: the first read has N's (rejected), the second is 7 bp"""
    fq = H.write_fastq(tmp_path / "short.fastq", golden["too_short_seqs"], qual_len=100)
    with pytest.raises(O.OracleError) as e:
        O.OracleGIR.create(40, [fq], "fastq", False)
    assert e.value.code == O.KO_ERR_SHORT_READ


def test_codec_known_answers(golden):
    kat = golden["codec_kat"]
    for s, expect in kat["compress_edge"]:
        assert list(O.compress_edge(s.encode())) == expect
        assert O.decompress_edge(bytes(expect)) == s.encode()
    enc = kat["encode_fasta_symbol"]
    block = 0
    for c in "ACGT":
        assert O.encode_fasta_symbol(ord(c), 0) == enc[c]
        block = O.encode_fasta_symbol(ord(c), block)
    assert block == enc["block_ACGT"]
    for by, expect in kat["shift_right"]["by"]:
        assert list(O.shift_right_bit_array(bytes(kat["shift_right"]["input"]), by)) == expect
    for inp, rem, out, s_in, s_out in kat["reverse_compressed_node"]:
        assert list(O.reverse_compressed_node(bytes(inp), rem)) == out
        assert list(O.reverse_compressed_node(bytes(out), rem)) == inp  # involution, compress.rs:629
        assert list(O.compress_node(s_in.encode())) == inp
        assert list(O.compress_node(s_out.encode())) == out
        assert H.revcomp(s_in) == s_out


def test_standardize_known_answer(golden):
    """standardizer.rs:254-279: (G,k,t)=(17,3,3) keeps 13 edges, weights 2 at idx 3,4 else 1"""
    import ctypes as C
    kat = golden["standardize_kat"]
    w = np.array(kat["weights"], np.uint32)
    err = C.c_int(0)
    m = O.lib().ko_standardize_weights(w.ctypes.data, len(w), kat["G"], kat["k"], kat["t"], C.byref(err))
    assert err.value == 0 and m == len(kat["expect"])
    assert w[:m].tolist() == kat["expect"]
    # ratio (10,0,10,0) -> 1.0 (standardizer.rs:138-141): weights unchanged
    w = np.array([4, 6], np.uint32)
    m = O.lib().ko_standardize_weights(w.ctypes.data, 2, 10, 0, 0, C.byref(err))
    assert err.value == 0 and m == 2 and w.tolist() == [4, 6]


@pytest.mark.parametrize("k", [3, 4, 5, 6, 21, 31, 32, 33, 40, 63, 64])
def test_byte_level_revcomp_equals_string_revcomp(k):
    """compress_kmer_with_rev_compl (compress.rs:34-48) == packing the reverse-complement string"""
    rng = np.random.default_rng(k)
    for _ in range(50):
        s = "".join(rng.choice(list("ACGT"), size=k))
        fw, rv = O.compress_kmer_with_rev_compl(s.encode())
        assert fw == O.compress_kmer(s.encode())
        assert rv == O.compress_kmer(H.revcomp(s).encode())


@pytest.mark.parametrize("k,rc", [(5, False), (5, True), (6, True), (31, True), (32, False), (33, True), (40, True),
                                  (64, True)])
def test_oracle_equals_python_set_level_twin(k, rc):
    """C oracle (node-keyed, chained) vs the 10-line set-level restatement of SURVEY Appendix A"""
    rng = np.random.default_rng(100 * k + rc)
    seqs = H.random_reads(rng, 60, k, k + 90, n_rate=0.15)
    edges, reads, nbytes = H.py_build(seqs, k, rc)
    g = O.OracleGIR(k)
    g.add_reads(*H.batch_of(seqs), rc)
    assert (g.accepted_reads, g.accepted_bytes) == (reads, nbytes)
    hi, lo, w = g.export_edges()
    ehi, elo, ew = H.py_sorted_arrays(edges)
    assert np.array_equal(hi, ehi) and np.array_equal(lo, elo) and np.array_equal(w, ew)
    assert g.counts() == (len(H.py_nodes(edges)), len(edges))
    # filter, then the orphan-node rule (pruner.rs:95-119): N' = prefixes/suffixes of E'
    g.remove_weak_edges(2)
    kept = {e: x for e, x in edges.items() if x >= 2}
    assert g.counts() == (len(H.py_nodes(kept)), len(kept))


def test_oracle_derived_goldens_are_reproducible(golden, tmp_path):
    """the committed oracle-derived numbers (rc=true, other k) still come out of the oracle"""
    for key in ("data1/k40/rc1", "data2/k40/rc1", "data3/k31/rc1", "data3/k63/rc0"):
        name, ks, rcs = key.split("/")
        fq = H.write_fastq(tmp_path / "g.fastq", H.golden_seqs(name))
        g, nbytes = O.OracleGIR.create(int(ks[1:]), [fq], "fastq", rcs == "rc1")
        exp = golden["oracle_derived"][key]
        assert nbytes == exp["bytes"] and list(g.counts()) == [exp["nodes"], exp["edges"]]
        assert list(g.digest()) == exp["digest"]
    # SURVEY Appendix B (derived, rc=true, k=40)
    d = golden["oracle_derived"]
    assert (d["data1/k40/rc1"]["nodes"], d["data1/k40/rc1"]["edges"]) == (124, 122)
    assert (d["data2/k40/rc1"]["nodes"], d["data2/k40/rc1"]["edges"]) == (11376, 11194)
    assert (d["data3/k40/rc1"]["nodes"], d["data3/k40/rc1"]["edges"]) == (28892, 28426)


def test_fasta_reader_and_multi_file(tmp_path):
    seqs = H.golden_seqs("data2")[:20]
    fa = H.write_fasta(tmp_path / "a.fasta", seqs)
    fq = H.write_fastq(tmp_path / "a.fastq", seqs)
    ga, ba = O.OracleGIR.create(40, [fa], "fasta", True)
    gq, bq = O.OracleGIR.create(40, [fq], "fastq", True)
    assert ba == bq and ga.digest() == gq.digest()
    fq2 = H.write_fastq(tmp_path / "b.fastq", H.golden_seqs("data2")[20:40])
    g2, b2 = O.OracleGIR.create(40, [fq, fq2], "fastq", True)
    g3, b3 = O.OracleGIR.create(40, [H.write_fastq(tmp_path / "c.fastq", H.golden_seqs("data2")[:40])], "fastq", True)
    assert b2 == b3 and g2.digest() == g3.digest()
    with pytest.raises(O.OracleError):
        O.OracleGIR.create(40, [str(tmp_path / "missing.fastq")])


def test_synthetic_reads_are_deterministic_and_error_rate_is_right():
    a = O.synth_reads(0x6B61746F6D65 + 1, 100000, 100, 5000, 10, 2010)
    b = O.synth_reads(0x6B61746F6D65 + 1, 100000, 100, 5000, 10, 2010)
    assert np.array_equal(a, b) and set(np.unique(a)) <= set(b"ACGT")
    clean = O.synth_reads(0x6B61746F6D65 + 1, 100000, 100, 0, 10, 2010)
    rate = float(np.mean(a != clean))
    assert 0.003 < rate < 0.007
    # an error-free forward-strand read is a slice of the genome
    genome = O.synth_genome(0x6B61746F6D65 + 1, 0, 100000).tobytes().decode()
    r0 = clean[:100].tobytes().decode()
    assert r0 in genome or H.revcomp(r0) in genome


# ------------------------------------------------------------------ BFCounter input
def _bfc_lines(rng, k, n):
    """unique canonical k-mers with counts, the shape of BFCounter's output"""
    from tests.helpers import kmer_int, revcomp
    seen, out = set(), []
    while len(out) < n:
        s = "".join(rng.choice(list("ACGT"), size=k))
        c = min(s, revcomp(s), key=kmer_int)
        if c not in seen:
            seen.add(c)
            out.append((c, int(rng.integers(1, 40))))
    return out


def test_bfcounter_input_matches_the_set_level_twin(tmp_path):
    """create_bfc (builder.rs:79-115) + add_read_bfc (pt_graph.rs:318-329) restated on the GIR:
    every line is one edge of its count, plus the reverse complement with rc; counts below the
    threshold are skipped before `total += len`.  The reference holds no fixture for this input
    (parity unpinned by reference tests); the witness is the set-level twin."""
    import numpy as np
    from collections import Counter
    from oracle import oracle as O
    from tests.helpers import kmer_int, py_nodes, revcomp
    rng = np.random.default_rng(11)
    for k, rc, t in ((31, True, 0), (40, False, 3), (6, True, 2), (63, True, 5)):
        lines = _bfc_lines(rng, k, 300)
        if k % 2 == 0 and rc:
            lines.append(("ACG" * (k // 6) + "CGT" * (k // 6), 7))  # a palindrome: both strands are one edge
            assert revcomp(lines[-1][0]) == lines[-1][0]
        path = tmp_path / f"bfc_{k}.txt"
        path.write_text("".join(f"{s}\t{w}\n" for s, w in lines))
        g, total = O.OracleGIR.create_bfc(k, [str(path)], rc, t)
        want = Counter()
        for s, w in lines:
            if w < t:
                continue
            want[s] += w
            if rc:
                want[revcomp(s)] += w
        assert total == sum(len(s) for s, w in lines if w >= t)
        hi, lo, wt = g.export_edges()
        got = {(int(h) << 64) | int(l): int(x) for h, l, x in zip(hi, lo, wt)}
        assert got == {kmer_int(s): w & 0xFFFFFFFF for s, w in want.items()}
        assert g.counts() == (len(py_nodes(want)), len(want))
    # errors: short k-mer, missing count, unparsable count, missing file
    for body, code in (("ACGT\t3\n", O.KO_ERR_SHORT_READ), ("A" * 31 + "\n", O.KO_ERR_BAD_RECORD),
                       ("A" * 31 + "\tx\n", O.KO_ERR_BAD_RECORD), ("A" * 30 + "N\t2\n", O.KO_ERR_BAD_RECORD)):
        p = tmp_path / "bad.txt"
        p.write_text(body)
        with pytest.raises(O.OracleError) as e:
            O.OracleGIR.create_bfc(31, [str(p)], True, 0)
        assert e.value.code == code
    with pytest.raises(O.OracleError):
        O.OracleGIR.create_bfc(31, [str(tmp_path / "missing.txt")], True, 0)


@pytest.mark.parametrize("k,rc", [(4, True), (5, True), (31, True), (31, False), (32, True), (32, False), (33, True),
                                  (40, True), (63, True), (64, True), (64, False)])
def test_optimistic_cpu_counter_equals_the_faithful_oracle(k, rc):
    """katome_oracle_mt.c (rolling, canonical, edge keyed, hash-sharded over threads: the "optimistic CPU"
    line of the bench) produces the digest of the faithful restatement, whatever the thread count."""
    rng = np.random.default_rng(7 * k + rc)
    seqs = H.random_reads(rng, 120, k, k + 90, n_rate=0.1)
    seqs += ["T" * (k + 7), "A" * (k + 3), "AT" * k, "ACGT" * k]  # all-ones key, palindromes (even k)
    bases, offsets = H.batch_of(seqs)
    g = O.OracleGIR(k)
    g.add_reads(bases, offsets, rc)
    for threads in (1, 2, 5):
        d, nr, nb = O.mt_build_digest(k, bases, offsets, rc, threads)
        assert d == g.digest()
        assert (nr, nb) == (g.accepted_reads, g.accepted_bytes)
    # a read shorter than k voids the build in both (hm_gir.rs:40)
    bases, offsets = H.batch_of(seqs + ["ACG"[: min(3, k - 1)]])
    with pytest.raises(O.OracleError) as e:
        O.mt_build_digest(k, bases, offsets, rc, 3)
    assert e.value.code == O.KO_ERR_SHORT_READ


@pytest.mark.parametrize("k,rc", [(4, True), (6, True), (31, True), (31, False), (32, False), (33, True), (40, False),
                                  (63, True), (64, True), (64, False)])
def test_mt_counter_object_follows_the_faithful_oracle_through_filter_and_standardize(k, rc):
    """MtCounter (the checker of the BASELINE-size GPU builds) == OracleGIR at every stage: batches added one by
    one (tables grow with the distinct keys), remove_weak_edges (edges.rs:51-58), more reads after the filter,
    standardize_edges (standardizer.rs:42-70), and the degenerate ratio."""
    rng = np.random.default_rng(11 * k + rc)
    genome = "".join(rng.choice(list("ACGT"), 3000))
    batches = [H.random_reads(rng, 400, k, k + 60, genome=genome, n_rate=0.05) for _ in range(3)]
    batches[1] += ["T" * (k + 9), "AT" * k, "ACGT" * k, "A" * k]
    g = O.OracleGIR(k)
    for threads in (1, 3, 8):
        g = O.OracleGIR(k)
        m = O.MtCounter(k, rc, threads)
        for seqs in batches[:2]:
            bases, offsets = H.batch_of(seqs)
            g.add_reads(bases, offsets, rc)
            m.add_reads(bases, offsets)
            assert m.digest() == g.digest()
        assert m.counters() == (g.accepted_reads, g.accepted_bytes)
        g.remove_weak_edges(3)
        m.remove_weak_edges(3)
        assert m.digest() == g.digest() and g.digest()[1] > 0
        bases, offsets = H.batch_of(batches[2])  # a removed edge that comes back starts from weight 0 again
        g.add_reads(bases, offsets, rc)
        m.add_reads(bases, offsets)
        assert m.digest() == g.digest()
        g.standardize_edges(2 * len(genome), k, 2)
        m.standardize_edges(2 * len(genome), k, 2)
        assert m.digest() == g.digest() and g.digest()[1] > 0
        g.remove_weak_edges(1 << 30)
        m.remove_weak_edges(1 << 30)
        assert m.digest() == g.digest() == (0, 0, 0, 0)
        with pytest.raises(O.OracleError) as e:
            m.standardize_edges(1000, k, 1)
        assert e.value.code == O.KO_ERR_DEGENERATE


def test_synth_reads_mt_writes_the_same_bytes():
    a = O.synth_reads(0x6B61746F6D65 + 1, 50_000, 100, 5000, 17, 1017)
    for threads in (1, 3, 7):
        assert np.array_equal(O.synth_reads_mt(0x6B61746F6D65 + 1, 50_000, 100, 5000, 17, 1017, threads), a)
