"""Parity of the CUDA path (through the C ABI) with the CPU oracle.  Bit-exact: identical
edge sets, weights, node counts and degree statistics, also after filtering/standardizing."""
import numpy as np
import pytest

from tests import helpers as H

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def K():
    import katome_b200
    from katome_b200 import _lib
    assert _lib.lib().ktg_device_count() > 0, "GPU tests need a CUDA device"
    return katome_b200


def _oracle(seqs, k, rc):
    from oracle import oracle as O
    g = O.OracleGIR(k)
    g.add_reads(*H.batch_of(seqs), rc)
    return g


def _assert_same(gpu, cpu, full_stats=True):
    assert gpu.digest() == cpu.digest()
    hi, lo, w = gpu.export_edges(sorted=True)
    ehi, elo, ew = cpu.export_edges()
    assert np.array_equal(lo, elo) and np.array_equal(hi, ehi) and np.array_equal(w, ew)
    assert gpu.counts() == cpu.counts()
    if full_stats:
        a, b = gpu.collection_stats(), cpu.collection_stats()
        for key in b:
            if isinstance(b[key], float):
                assert (np.isnan(a[key]) and np.isnan(b[key])) or a[key] == b[key], key
            else:
                assert a[key] == b[key], key


# ------------------------------------------------------------------ C1: the reference's fixtures
@pytest.mark.parametrize("name", ["data1", "data2", "data3"])
def test_reference_fixtures_k40_norc(K, golden, tmp_path, name):
    """tests/build.rs:27-28,46-89 through Build::create on the GPU"""
    ref = golden["reference_pinned"][name]
    fq = H.write_fastq(tmp_path / f"{name}.fastq", H.golden_seqs(name))
    g, nbytes = K.GpuGIR.create([fq], "fastq", False, 0, k=40)
    assert nbytes == ref["bytes"]
    assert g.counts() == (ref["nodes"], ref["edges"])
    st = g.collection_stats()
    for key in ("max_edge_weight", "max_in_degree", "max_out_degree", "incoming_vert_count", "outgoing_vert_count"):
        assert st[key] == ref[key], key
    assert round(st["avg_edge_weight"], 2) == ref["avg_edge_weight"]
    assert round(st["avg_out_degree"], 2) == ref["avg_out_degree"]


def test_reference_filter_cases(K, golden, tmp_path):
    """tests/pruner.rs:240-251, 260-262, 272"""
    for case in golden["reference_pinned_filter"]:
        fq = H.write_fastq(tmp_path / "f.fastq", H.golden_seqs(case["file"]))
        g, _ = K.GpuGIR.create([fq], "fastq", False, 0, k=40)
        g.remove_weak_edges(case["threshold"])
        assert g.counts() == (case["nodes"], case["edges"]), case


@pytest.mark.parametrize("k", [21, 31, 32, 33, 40, 63, 64])
@pytest.mark.parametrize("rc", [False, True])
def test_fixtures_against_committed_goldens(K, golden, tmp_path, k, rc):
    for name in ("data1", "data2", "data3"):
        exp = golden["oracle_derived"][f"{name}/k{k}/rc{int(rc)}"]
        fq = H.write_fastq(tmp_path / f"{name}.fastq", H.golden_seqs(name))
        g, nbytes = K.GpuGIR.create([fq], "fastq", rc, 0, k=k)
        assert nbytes == exp["bytes"]
        assert list(g.digest()) == exp["digest"]
        assert list(g.counts()) == [exp["nodes"], exp["edges"]]
        st = g.collection_stats()
        for key, v in exp["stats"].items():
            assert st[key] == v, (name, key)
        for t in (2, 3):
            g2, _ = K.GpuGIR.create([fq], "fastq", rc, 0, k=k)
            g2.remove_weak_edges(t)
            assert list(g2.digest()) == exp[f"filter_t{t}"]["digest"]
            assert list(g2.counts()) == exp[f"filter_t{t}"]["counts"]
            g2.close()
        g.close()


def test_too_short_read_voids_the_build(K, golden, tmp_path):
    fq = H.write_fastq(tmp_path / "short.fastq", golden["too_short_seqs"], qual_len=100)
    with pytest.raises(K.ReadTooShort):
        K.GpuGIR.create([fq], "fastq", False, 0, k=40)
    g = K.GpuGIR(40, True)
    with pytest.raises(K.ReadTooShort):
        g.add_read_fastaq(b"ACGTACG")
    with pytest.raises(K.KatomeError):  # the build stays void
        g.add_read_fastaq(b"A" * 50)


def test_file_errors_mirror_check_files(K, tmp_path):
    with pytest.raises(K.KatomeError) as e:
        K.GpuGIR.create([str(tmp_path / "nope.fastq")], "fastq", True, 0, k=31)
    assert "resolve path" in str(e.value)
    with pytest.raises(K.KatomeError) as e:
        K.GpuGIR.create([str(tmp_path)], "fastq", True, 0, k=31)
    assert "is a directory" in str(e.value)
    bad = tmp_path / "bad.fastq"
    bad.write_text("ACGT\nACGT\n+\nIIII\n")
    with pytest.raises(K.KatomeError):
        K.GpuGIR.create([str(bad)], "fastq", True, 0, k=3)
    for k in (1, 2, 65):
        with pytest.raises(K.KatomeError):
            K.GpuGIR(k)


def test_fasta_and_multiple_files(K, tmp_path):
    seqs = H.golden_seqs("data2")
    fa = H.write_fasta(tmp_path / "a.fasta", seqs[:50])
    fq1 = H.write_fastq(tmp_path / "a.fastq", seqs[:50])
    fq2 = H.write_fastq(tmp_path / "b.fastq", seqs[50:])
    ga, ba = K.GpuGIR.create([fa], "fasta", True, 0, k=40)
    gq, bq = K.GpuGIR.create([fq1], "fastq", True, 0, k=40)
    assert ba == bq and ga.digest() == gq.digest()
    g2, b2 = K.GpuGIR.create([fq1, fq2], "fastq", True, 0, k=40)
    cpu = _oracle(seqs, 40, True)
    assert b2 == cpu.accepted_bytes
    _assert_same(g2, cpu)


# ------------------------------------------------------------------ random / ragged / edge cases
@pytest.mark.parametrize("k", [3, 4, 5, 16, 17, 31, 32, 33, 34, 40, 47, 48, 63, 64])
@pytest.mark.parametrize("rc", [False, True])
def test_random_ragged_reads(K, k, rc):
    rng = np.random.default_rng(1000 * k + rc)
    seqs = H.random_reads(rng, 300, k, k + 200, n_rate=0.1)
    seqs += ["A" * (k + 5), "T" * (k + 40), "ACGT" * 40, "AT" * 70, "", "N"]  # homopolymers, palindromes
    seqs = [s for s in seqs if len(s) >= k or any(c not in "ACGT" for c in s)]  # short valid reads abort
    cpu = _oracle(seqs, k, rc)
    g = K.GpuGIR(k, rc)
    nr, nb = g.add_reads(*H.batch_of(seqs))
    assert (nr, nb) == (cpu.accepted_reads, cpu.accepted_bytes)
    _assert_same(g, cpu)
    # python set-level twin as a second witness
    edges, _, _ = H.py_build(seqs, k, rc)
    assert g.counts() == (len(H.py_nodes(edges)), len(edges))
    for t in (2, 3):
        cpu.remove_weak_edges(t)
        g.remove_weak_edges(t)
        _assert_same(g, cpu)
    g.close()


@pytest.mark.parametrize("k", [31, 40])
def test_direct_and_partitioned_paths_agree(K, k):
    rng = np.random.default_rng(7 + k)
    genome = "".join(rng.choice(list("ACGT"), size=20000))
    seqs = H.random_reads(rng, 3000, 100, 150, genome=genome, n_rate=0.02)
    cpu = _oracle(seqs, k, True)
    for kw in ({"force_direct": True}, {"force_partition": True}, {"force_partition": True, "sub_table_log2_bytes": 16}):
        g = K.GpuGIR(k, True, **kw)
        g.add_reads(*H.batch_of(seqs))
        _assert_same(g, cpu)
        assert g.info()["partitioned"] == int("force_partition" in kw)
        g.close()


@pytest.mark.parametrize("k", [4, 21, 31, 32, 33, 40, 63, 64])
@pytest.mark.parametrize("rc", [False, True])
def test_paged_update_path(K, k, rc):
    """two-level partition + streaming page update (shared-memory inserts), several batches,
    fresh and non-fresh tables, palindromes (even k) and the all-T key at full width"""
    rng = np.random.default_rng(31 * k + rc)
    genome = "".join(rng.choice(list("ACGT"), size=30000))
    seqs = H.random_reads(rng, 4000, max(k, 60), 160, genome=genome, n_rate=0.02)
    seqs += ["T" * (k + 20), "A" * (k + 7), "AT" * 60, "ACGT" * 40]
    cpu = _oracle(seqs, k, rc)
    import os
    for kw, factor in (({}, None), ({"sub_table_log2_bytes": 18}, 1), ({"edges_count": 2 * 30000 * 3}, 1),
                       ({"options": {"page_nbuf": 2}}, 1), ({"options": {"page_nbuf": 2, "page_threads": 1024}}, None)):
        # stage_factor_milli = 1: flush the staged buckets after every batch instead of once at the end
        g = K.GpuGIR(k, rc, force_pages=True, **kw)
        if factor:
            g.set_option("stage_factor_milli", factor)
        third = len(seqs) // 3
        for part in (seqs[:third], seqs[third:2 * third], seqs[2 * third:]):
            g.add_reads(*H.batch_of(part))
        _assert_same(g, cpu)
        info = g.info()
        assert info["page_updates"] == (3 if factor else 1) and info["partitioned"] == 1
        g.reset()
        g.add_reads(*H.batch_of(seqs))
        _assert_same(g, cpu, full_stats=False)
        g.close()


@pytest.mark.parametrize("k,rc", [(31, True), (32, False), (63, True)])
def test_heavy_hitters_overflow_the_page_buckets_without_loss(K, k, rc):
    """Skewed input on the page path: 10 % homopolymer and dinucleotide reads put millions of windows on a
    handful of canonical k-mers, far beyond what their pages' buckets hold (1.125 x mean + 1024).  The
    overflow goes to the stage's spill list, which is as large as the stage, so the build neither fails
    nor loses a key (the reference builds such inputs: hm_gir.rs:39-87 has no capacity anywhere)."""
    from oracle import oracle as O
    rng = np.random.default_rng(5 * k + rc)
    genome = "".join(rng.choice(list("ACGT"), size=200_000))
    n = 60_000
    seqs = [genome[i:i + 100] for i in rng.integers(0, len(genome) - 100, size=n)]
    for i in range(0, n, 10):
        seqs[i] = ("A" * 100, "T" * 100, "AC" * 50, "G" * 100, "ACGT" * 25)[(i // 10) % 5]
    bases, offsets = H.batch_of(seqs)
    m = O.MtCounter(k, rc)
    m.add_reads(bases, offsets)
    for kw in ({"force_pages": True, "sub_table_log2_bytes": 20, "edges_count": 900_000},
               {"force_pages": True, "sub_table_log2_bytes": 20, "edges_count": 900_000, "options": {"chunk_mb": 1}},
               {"force_pages": True}):
        g = K.GpuGIR(k, rc, **kw)
        assert g.add_reads(bases, offsets) == m.counters()
        assert g.digest() == m.digest(), kw
        assert g.info()["page_updates"] >= 1
        g.close()
    assert m.digest()[3] >= 1_200 * (100 - k + 1)  # one homopolymer k-mer alone: far more than a page bucket holds


@pytest.mark.parametrize("k,rc", [(21, True), (31, True), (32, False), (33, True), (40, False), (63, True), (64, False)])
def test_big_tile_level1_scatter(K, k, rc):
    """The level-1 scatter with the big tile (two extraction passes, 64 KB of keys per tile; the default when
    a table has more than 256 sub-tables, forced here with the l1_big option on tables of many small
    sub-tables): uniform and ragged reads with N / lowercase, T...T at full key width, and homopolymer
    reads that overflow their level-1 buckets into the spill list -- against the oracle, and bin for bin
    against the one-pass kernel."""
    from oracle import oracle as O
    rng = np.random.default_rng(77 * k + rc)
    genome = "".join(rng.choice(list("ACGT"), size=150_000))
    uniform = [genome[i:i + 100] for i in rng.integers(0, len(genome) - 100, size=40_000)]
    uniform[5] = uniform[5][:50] + "N" + uniform[5][51:]
    uniform[6] = uniform[6].lower()
    for i in range(0, len(uniform), 9):  # 11 % low-complexity reads: a few bins get far more than their bucket holds
        uniform[i] = ("T" * 100, "A" * 100, "GT" * 50, "ACGT" * 25)[(i // 9) % 4]
    ragged = H.random_reads(rng, 20_000, max(k, 66), 180, genome=genome, n_rate=0.02) + ["T" * (k + 3), "A" * k]
    for name, seqs in (("uniform", uniform), ("ragged", ragged)):
        bases, offsets = H.batch_of(seqs)
        m = O.MtCounter(k, rc)
        m.add_reads(bases, offsets)
        got = []
        for big in (1, 0):
            for kw in ({"force_pages": True, "sub_table_log2_bytes": 16, "edges_count": 600_000},   # hundreds of sub-tables
                       {"force_partition": True, "no_pages": True, "sub_table_log2_bytes": 17, "edges_count": 600_000},
                       {"force_pages": True, "sub_table_log2_bytes": 18, "edges_count": 600_000, "options": {"chunk_mb": 1}}):
                kw = dict(kw)
                kw["options"] = dict(kw.get("options", {}), l1_big=big)
                g = K.GpuGIR(k, rc, **kw)
                assert g.add_reads(bases, offsets) == m.counters(), (name, kw)
                assert g.digest() == m.digest(), (name, kw)
                got.append(g.counts())
                g.close()
        assert len(set(got)) == 1, got


def test_paged_and_atomic_paths_agree_at_scale(K):
    """mini-C2 shape, default path selection (page update) against L2 atomics only"""
    from katome_b200.workloads import Workload
    wl = Workload("mini-C2", 1, 1_000_000, 100, 40, 5000, 31)
    n, L = wl.n_reads, wl.read_len
    d = torch.empty(n * L, dtype=torch.uint8, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    K.synth_reads_device(d, wl.seed, wl.genome_len, L, wl.err_ppm, 0, n, stream=s)
    offs = torch.arange(0, (n + 1) * L, L, dtype=torch.int64, device="cuda")
    res = []
    for kw in ({}, {"no_pages": True}, {"sub_table_log2_bytes": 20}):
        g = K.GpuGIR(31, True, stream=s, edges_count=wl.expected_distinct_edges(), **kw)
        g.add_reads_device(d, offs, n, n * L)
        res.append((g.digest(), g.counts()))
        assert g.info()["page_updates"] == (0 if "no_pages" in kw else 1)
        assert g.digest()[2] == 2 * wl.n_windows
        g.close()
    assert res[0] == res[1] == res[2]
    # k = 63: u128 keys, 128-bit shared-memory CAS
    res = []
    for kw in ({}, {"no_pages": True}):
        g = K.GpuGIR(63, True, stream=s, **kw)
        g.add_reads_device(d, offs, n, n * L)
        res.append((g.digest(), g.counts()))
        g.close()
    assert res[0] == res[1] and res[0][0][2] == 2 * n * (L - 63 + 1)


@pytest.mark.parametrize("k,rc", [(31, True), (32, False), (40, True), (63, True)])
def test_host_batcher_large_call(K, k, rc):
    """ktg_add_reads on a call of many chunks (option chunk_mb=1 makes a 5 MB batch "large"): tapered last
    chunks, held-back flush with one flush on the way, and the eager page stage (every chunk's keys are
    moved on to page buckets at once, a flush is only the page sweep) -- against the oracle, with the
    eager stage off, with a table that has to grow at the flush (page buckets of a geometry that is
    gone fall back to L2 atomics), with ragged reads, and over two calls into one stage."""
    import os
    rng = np.random.default_rng(1000 + k)
    genome = "".join(rng.choice(list("ACGT"), size=60000))
    uniform = [genome[i:i + 100] for i in rng.integers(0, len(genome) - 100, size=50000)]
    uniform[17] = uniform[17][:40] + "N" + uniform[17][41:]
    uniform += ["T" * 100, "A" * 100, "AT" * 50, "ACGT" * 25]
    ragged = H.random_reads(rng, 30000, max(k, 70), 190, genome=genome, n_rate=0.02)
    cases = []
    for name, seqs in (("uniform", uniform), ("ragged", ragged)):
        cases.append((name, seqs, _oracle(seqs, k, rc)))
    distinct = 2 * len(genome) * 2
    try:
        for name, seqs, cpu in cases:
            bases, offsets = H.batch_of(seqs)
            for kw, env in (({"force_pages": True, "sub_table_log2_bytes": 18, "edges_count": distinct}, {}),
                            ({"force_pages": True, "sub_table_log2_bytes": 18, "edges_count": distinct}, {"eager_pages": 0}),
                            ({"force_pages": True, "sub_table_log2_bytes": 18, "edges_count": distinct}, {"flush_pct": 30, "stage_bufs": 3}),
                            ({"force_pages": True, "sub_table_log2_bytes": 16, "edges_count": 3000}, {}),  # must grow
                            ({"force_pages": True, "sub_table_log2_bytes": 16}, {})):                     # no hint at all
                try:
                    g = K.GpuGIR(k, rc, profile=True, options=dict(env, chunk_mb=1), **kw)
                    nr, nb = g.add_reads(bases, offsets)
                    assert (nr, nb) == (cpu.accepted_reads, cpu.accepted_bytes), (name, kw, env)
                    _assert_same(g, cpu, full_stats=False)
                    if kw.get("edges_count") == distinct:
                        sweeps, scatters = g.info()["page_updates"], g.profile()["scatter_pages"]["launches"]
                        assert sweeps >= 2, (name, g.info())  # the flush on the way + the final one
                        # eager: one level-2 scatter per chunk; otherwise one per sweep
                        assert (scatters == sweeps) if "eager_pages" in env else (scatters >= 4 > sweeps), (scatters, sweeps)
                    g.close()
                finally:
                    pass
        # two large calls into one builder; the second finds the first one's stage open
        name, seqs, cpu2 = cases[0][0], cases[0][1] + cases[1][1], None
        cpu2 = _oracle(seqs, k, rc)
        g = K.GpuGIR(k, rc, force_pages=True, sub_table_log2_bytes=18, edges_count=distinct, options={"chunk_mb": 1})
        g.add_reads(*H.batch_of(cases[0][1]))
        g.add_reads(*H.batch_of(cases[1][1]))
        _assert_same(g, cpu2, full_stats=False)
        g.close()
    finally:
        pass


@pytest.mark.parametrize("k,rc", [(4, True), (31, True), (32, False), (33, True), (34, False), (40, True), (63, True), (64, True)])
def test_graph_export_for_convert(K, k, rc):
    """the hand-off to Convert::create_from -- sorted node set, edges with the indices
    of their prefix / suffix nodes, weights, and the edges in compress_edge bytes (compress.rs:250-271)"""
    from oracle import oracle as O
    rng = np.random.default_rng(5 * k + rc)
    genome = "".join(rng.choice(list("ACGT"), size=4000))
    seqs = H.random_reads(rng, 600, max(k, 50), 140, genome=genome, n_rate=0.02) + ["T" * (k + 9), "ACGT" * 30]
    cpu = _oracle(seqs, k, rc)
    g = K.GpuGIR(k, rc)
    g.add_reads(*H.batch_of(seqs))
    out = g.export_graph()
    nhi, nlo = cpu.export_nodes()
    ehi, elo, ew = cpu.export_edges()
    assert np.array_equal(out["node_hi"], nhi) and np.array_equal(out["node_lo"], nlo)
    assert np.array_equal(out["weight"], ew)
    nodes = [(int(h) << 64) | int(l) for h, l in zip(nhi.tolist(), nlo.tolist())]
    index = {v: i for i, v in enumerate(nodes)}
    mask = (1 << (2 * (k - 1))) - 1
    for e, (h, l) in enumerate(zip(ehi.tolist(), elo.tolist())):
        edge = (h << 64) | l
        assert out["src"][e] == index[edge >> 2] and out["dst"][e] == index[edge & mask]
        if e % 7 == 0:  # the byte string of every 7th edge against the oracle's compress_edge
            text = "".join("ACGT"[(edge >> (2 * (k - 1 - j))) & 3] for j in range(k)).encode()
            assert bytes(out["edge_bytes"][e]) == O.compress_edge(text)
    assert O.compress_edge(b"AGGTCG") == bytes([2, 0b00101011, 0b01100000])  # compress.rs:244-248
    g.close()


@pytest.mark.parametrize("k,rc,shards", [(5, True, 0), (31, True, 0), (31, False, 3), (40, False, 0), (63, True, 2)])
def test_externals_seed_the_dead_path_search(K, k, rc, shards):
    """ktg_export_externals == the `Externals` iterator of remove_dead_paths (pruner.rs:165-195) over the graph
    Convert::create_from would build from the oracle's edges: node index ascending (sorted nodes), Input when
    the node has no incoming edge, else Output when it has no outgoing edge.  Also after the filter, and on a
    sharded handle."""
    rng = np.random.default_rng(9 * k + rc)
    genome = "".join(rng.choice(list("ACGT"), size=3000))
    seqs = H.random_reads(rng, 300, max(k, 40), 120, genome=genome, n_rate=0.02) + ["ACGT" * 30, "T" * (k + 4)]
    cpu = _oracle(seqs, k, rc)
    g = K.GpuGIR(k, rc, **({"device_ids": [0] * shards} if shards else {}))
    g.add_reads(*H.batch_of(seqs))

    def twin(cpu):
        ehi, elo, _ = cpu.export_edges()
        edges = [(int(h) << 64) | int(l) for h, l in zip(ehi.tolist(), elo.tolist())]
        mask = (1 << (2 * (k - 1))) - 1
        nodes = sorted({e >> 2 for e in edges} | {e & mask for e in edges})
        idx = {v: i for i, v in enumerate(nodes)}
        has_out, has_in = set(idx[e >> 2] for e in edges), set(idx[e & mask] for e in edges)
        ids, kinds = [], []
        for v in range(len(nodes)):
            if v not in has_in:
                ids.append(v), kinds.append(0)
            elif v not in has_out:
                ids.append(v), kinds.append(1)
        return np.array(ids, np.uint64), np.array(kinds, np.uint8)

    for t in (0, 2):
        if t:
            g.remove_weak_edges(t), cpu.remove_weak_edges(t)
        ids, kinds = g.export_externals()
        want_ids, want_kinds = twin(cpu)
        assert np.array_equal(ids, want_ids) and np.array_equal(kinds, want_kinds), (t, len(ids), len(want_ids))
        st = g.collection_stats()  # sources + the sinks that are not sources
        assert int((kinds == 0).sum()) == st["incoming_vert_count"]
        assert len(ids) > 0 or k == 5  # (the 4^4 nodes of k = 5 all have edges both ways)
    g.close()


def test_all_T_at_full_key_width_without_canonicalisation(K):
    """k=32 / k=64, reverse_complement=false: TTT...T equals the all-ones 'empty' marker"""
    for k in (32, 64):
        seqs = ["T" * (k + 10), "ACGT" * 30, "T" * k, "G" + "T" * (k + 3)]
        for rc in (False, True):
            cpu = _oracle(seqs, k, rc)
            g = K.GpuGIR(k, rc)
            g.add_reads(*H.batch_of(seqs))
            _assert_same(g, cpu)
            g.remove_weak_edges(3)
            cpu.remove_weak_edges(3)
            _assert_same(g, cpu)
            g.close()


def test_empty_and_minimal_inputs(K):
    g = K.GpuGIR(31, True)
    assert g.add_reads(np.zeros(0, np.uint8), np.zeros(1, np.uint64)) == (0, 0)
    assert g.counts() == (0, 0) and g.digest() == (0, 0, 0, 0)
    assert g.add_reads(*H.batch_of(["ACGTNACGT" * 10, "acgt" * 20])) == (0, 0)  # all rejected
    assert g.counts() == (0, 0)
    s = "ACGTTGCATGCATGCCGATAGCTAGCTAGGA"  # exactly k: one window
    assert g.add_reads(*H.batch_of([s])) == (1, 31)
    assert g.counts() == (4, 2)
    assert g.dump() == _oracle([s], 31, True).dump()
    g.close()


def test_batches_accumulate_and_table_grows_without_a_hint(K):
    from oracle import oracle as O
    G, L, k = 1_500_000, 100, 31
    n = 40_000
    reads = O.synth_reads(0xABCDEF, G, L, 5000, 0, n)
    offsets = (np.arange(n + 1, dtype=np.uint64) * L)
    cpu = O.OracleGIR(k)
    cpu.add_reads(reads, offsets, True)
    g = K.GpuGIR(k, True, sub_table_log2_bytes=16)  # starts at 1 Mi slots, must grow + partition
    step = 10_000
    for r0 in range(0, n, step):
        g.add_reads(reads[r0 * L:(r0 + step) * L], offsets[r0:r0 + step + 1] - offsets[r0])
    _assert_same(g, cpu, full_stats=True)
    info = g.info()
    assert info["grow_events"] >= 1 and info["occupied_slots"] * 2 == cpu.counts()[1]
    # one shot with a capacity hint gives the same table content
    g2 = K.GpuGIR(k, True, edges_count=cpu.counts()[1])
    g2.add_reads(reads, offsets)
    assert g2.digest() == cpu.digest() and g2.info()["grow_events"] == 0
    g2.reset()
    assert g2.counts() == (0, 0)
    g2.add_reads(reads, offsets)
    assert g2.digest() == cpu.digest()
    g.close(), g2.close()


@pytest.mark.parametrize("k,rc", [(31, True), (40, True), (40, False), (63, True)])
def test_standardize_edges(K, k, rc):
    rng = np.random.default_rng(k)
    genome = "".join(rng.choice(list("ACGT"), size=3000))
    seqs = H.random_reads(rng, 1500, 100, 150, genome=genome)
    for t in (0, 2, 5):
        cpu = _oracle(seqs, k, rc)
        g = K.GpuGIR(k, rc)
        g.add_reads(*H.batch_of(seqs))
        cpu.standardize_edges(len(genome), k, t)
        g.standardize_edges(len(genome), k, t)
        _assert_same(g, cpu)
        g.close()
    g = K.GpuGIR(k, rc)
    g.add_reads(*H.batch_of(seqs))
    with pytest.raises(K.KatomeError):
        g.standardize_edges(k - 1, k, 0)  # G < k: the reference underflows/panics


def test_device_resident_inputs_and_synthetic_generator(K):
    from oracle import oracle as O
    seed, G, L, n = 0x6B61746F6D65 + 1, 200_000, 100, 5000
    d = torch.empty(n * L, dtype=torch.uint8, device="cuda")
    K.synth_reads_device(d, seed, G, L, 5000, 100, 100 + n, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    host = O.synth_reads(seed, G, L, 5000, 100, 100 + n)
    assert np.array_equal(d.cpu().numpy(), host)
    offs = torch.arange(0, (n + 1) * L, L, dtype=torch.int64, device="cuda")
    for k in (31, 63):
        g = K.GpuGIR(k, True, stream=torch.cuda.current_stream().cuda_stream)
        assert g.add_reads_device(d, offs, n, n * L, want_counts=True) == (n, n * L)
        cpu = O.OracleGIR(k)
        cpu.add_reads(host, np.arange(n + 1, dtype=np.uint64) * L, True)
        _assert_same(g, cpu)
        g.close()


def test_two_rank_sharding_on_one_gpu(K):
    """world_size=2 emulated with two handles on one GPU: partition -> exchange -> insert; the
    union of the shards is the oracle's table and every key sits on its owner."""
    from oracle import oracle as O
    seed, G, L, n, k = 77, 150_000, 150, 4000, 31
    host = O.synth_reads(seed, G, L, 5000, 0, n)
    cpu = O.OracleGIR(k)
    cpu.add_reads(host, np.arange(n + 1, dtype=np.uint64) * L, True)
    W = 2
    gs = [K.GpuGIR(k, True, world_size=W, rank=r) for r in range(W)]
    half = n // W
    recv = [[] for _ in range(W)]
    for r in range(W):
        d = torch.from_numpy(host[r * half * L:(r + 1) * half * L]).cuda()
        offs = torch.arange(0, (half + 1) * L, L, dtype=torch.int64, device="cuda")
        ptr, counts = gs[r].partition_reads_device(d, offs, half, half * L)
        assert sum(counts) == half * (L - k + 1)
        gs[r].finalize()
        total = sum(counts)
        buf = torch.as_tensor(K.DeviceArray(ptr, total), device="cuda").clone()
        o = 0
        for dst in range(W):
            recv[dst].append(buf[o:o + counts[dst]].clone())
            o += counts[dst]
    parts = []
    for r in range(W):
        keys = torch.cat(recv[r])
        gs[r].insert_keys_device(keys, keys.numel())
        hi, lo, w = gs[r].export_edges(sorted=True)
        assert all(gs[r].owner_of(0, int(x)) == r for x in lo[:200].tolist())
        parts.append((lo, w))
    lo = np.concatenate([p[0] for p in parts])
    w = np.concatenate([p[1] for p in parts])
    order = np.argsort(lo, kind="stable")
    _, elo, ew = cpu.export_edges()
    assert np.array_equal(lo[order], elo) and np.array_equal(w[order], ew)
    d0, d1 = gs[0].digest(), gs[1].digest()
    cd = cpu.digest()
    assert ((d0[0] + d1[0]) & (2**64 - 1), d0[1] + d1[1], d0[2] + d1[2], max(d0[3], d1[3])) == cd


def _fused_exchange_emulated(K, gs, batches, k):
    """The protocol of katome_b200.dist.ShardedGIR._add_reads_fused with all ranks in one process on
    one GPU: peers' receive buffers are plain device pointers, collectives are torch ops."""
    W = len(gs)
    gmax = max(max(nb - n * (k - 1), 0) for (_, _, n, nb) in batches)
    prep = [g.mg_prepare(gmax) for g in gs]
    peers = [p[0] for p in prep]
    cap = prep[0][2]
    assert all(p[2] == cap for p in prep)
    curs = [torch.as_tensor(K.DeviceArray(g.mg_scatter_reads_device(d, o, n, nb, peers), W), device="cuda")
            for g, (d, o, n, nb) in zip(gs, batches)]
    regs = [torch.as_tensor(K.DeviceArray(*g.mg_sketch(), "<i4"), device="cuda") for g in gs]
    m = torch.stack(regs).max(dim=0).values
    for r in regs:
        r.copy_(m)
    for r, g in enumerate(gs):
        got = torch.stack([c[r] for c in curs])  # what the all-to-all delivers to rank r
        fill = (got - r * cap).clamp_(max=cap)
        ends = torch.arange(W, dtype=torch.int64, device="cuda") * cap + fill
        g.mg_insert_buckets(ends, int(fill.sum().item()))
        torch.cuda.synchronize()
    words = gs[0].key_words()
    routed = [[] for _ in range(W)]
    n_spilled = 0
    for g in gs:
        ptr, n = g.mg_spill()
        n_spilled += n
        if n:
            out, counts = g.partition_keys_device(ptr, n)
            buf = torch.as_tensor(K.DeviceArray(out, n * words), device="cuda").clone()
            o = 0
            for dst in range(W):
                routed[dst].append(buf[o:o + counts[dst] * words])
                o += counts[dst] * words
    for r, g in enumerate(gs):
        if routed[r]:
            keys = torch.cat(routed[r])
            g.mg_insert_spill(keys, keys.numel() // words)
            torch.cuda.synchronize()
    return n_spilled


@pytest.mark.parametrize("k,W", [(31, 2), (31, 3), (40, 2), (63, 4)])
def test_fused_exchange_ranks_emulated_on_one_gpu(K, k, W):
    """fused multi-GPU path: scatter into the owners' receive buckets, level-2 scatter + page update
    from them; two batches, a skewed one (bucket overflow -> spill list) and growth without a hint"""
    from oracle import oracle as O
    seed, G, L, n = 1234 + k, 120_000, 150, 3000 * W
    host = O.synth_reads(seed, G, L, 5000, 0, n)
    skew = np.frombuffer((b"A" * L) * (2500 * W), dtype=np.uint8)  # one key, 2500*W*(L-k+1) times
    cpu = O.OracleGIR(k)
    cpu.add_reads(host, np.arange(n + 1, dtype=np.uint64) * L, True)
    cpu.add_reads(skew, np.arange(2500 * W + 1, dtype=np.uint64) * L, True)
    s = torch.cuda.current_stream().cuda_stream
    gs = [K.GpuGIR(k, True, world_size=W, rank=r, force_pages=True, stream=s) for r in range(W)]
    spilled = 0
    for data, per in ((host, n // W), (skew, 2500)):
        batches = []
        for r in range(W):
            d = torch.from_numpy(data[r * per * L:(r + 1) * per * L].copy()).cuda()
            offs = torch.arange(0, (per + 1) * L, L, dtype=torch.int64, device="cuda")
            batches.append((d, offs, per, per * L))
        spilled += _fused_exchange_emulated(K, gs, batches, k)
    assert spilled > 0  # the skewed batch must have exercised the spill route
    digs = [g.digest() for g in gs]
    cd = cpu.digest()
    assert (sum(d[0] for d in digs) & (2**64 - 1), sum(d[1] for d in digs), sum(d[2] for d in digs),
            max(d[3] for d in digs)) == cd
    parts = [g.export_edges(sorted=True) for g in gs]
    for r, (hi, lo, w) in enumerate(parts):
        assert all(gs[r].owner_of(int(h), int(x)) == r for h, x in list(zip(hi.tolist(), lo.tolist()))[:100])
    hi = np.concatenate([p[0] for p in parts]); lo = np.concatenate([p[1] for p in parts])
    w = np.concatenate([p[2] for p in parts])
    order = np.lexsort((lo, hi))
    ehi, elo, ew = cpu.export_edges()
    assert np.array_equal(hi[order], ehi) and np.array_equal(lo[order], elo) and np.array_equal(w[order], ew)
    assert sum(g.info()["page_updates"] for g in gs) > 0
    for g in gs:
        g.close()


def _skm_exchange_emulated(K, gs, batches, k, claim=None):
    """The protocol of ShardedGIR._add_reads_skm with all ranks in one process on one GPU.  `claim`
    under-provisions the receive buckets on purpose (forces the spill route).  -> spilled records"""
    W = len(gs)
    gmax = claim or max(max(nb - n * (k - 1), 0) for (_, _, n, nb) in batches)
    prep = [g.mg_skm_prepare(gmax) for g in gs]
    peers = [p[0] for p in prep]
    cap = prep[0][2]
    assert all(p[2] == cap for p in prep)
    outs = [g.mg_skm_scatter_reads_device(d, o, n, nb, peers) for g, (d, o, n, nb) in zip(gs, batches)]
    curs = [torch.as_tensor(K.DeviceArray(c, W), device="cuda") for c, _ in outs]
    kcs = [torch.as_tensor(K.DeviceArray(kc, W), device="cuda") for _, kc in outs]
    torch.cuda.synchronize()
    for r, g in enumerate(gs):
        got = torch.stack([c[r] for c in curs])  # what the all-to-all delivers to rank r
        n_keys = int(sum(int(kc[r]) for kc in kcs))
        fill = (got - r * cap).clamp_(max=cap)
        ends = torch.arange(W, dtype=torch.int64, device="cuda") * cap + fill
        g.mg_skm_insert_buckets(ends, n_keys)
        torch.cuda.synchronize()
    routed = [[] for _ in range(W)]
    n_spilled = 0
    for g in gs:
        ptr, n = g.mg_skm_spill()
        n_spilled += n
        if n:
            out, counts = g.mg_skm_partition_records(ptr, n)
            assert sum(counts) == n
            buf = torch.as_tensor(K.DeviceArray(out, n * 2), device="cuda").clone()
            o = 0
            for dst in range(W):
                routed[dst].append(buf[o:o + counts[dst] * 2])
                o += counts[dst] * 2
    for r, g in enumerate(gs):
        if routed[r]:
            recs = torch.cat(routed[r])
            g.mg_skm_insert_records(recs, recs.numel() // 2)
            torch.cuda.synchronize()
    return n_spilled, [int(sum(int(kc[r]) for kc in kcs)) for r in range(W)]


def _assert_shards_equal_oracle(gs, cpu, owner):
    digs = [g.digest() for g in gs]
    cd = cpu.digest()
    assert (sum(d[0] for d in digs) & (2**64 - 1), sum(d[1] for d in digs), sum(d[2] for d in digs),
            max(d[3] for d in digs)) == cd
    parts = [g.export_edges(sorted=True) for g in gs]
    for r, (hi, lo, w) in enumerate(parts):
        assert all(owner(gs[r], int(h), int(x)) == r for h, x in list(zip(hi.tolist(), lo.tolist()))[:200])
    hi = np.concatenate([p[0] for p in parts]); lo = np.concatenate([p[1] for p in parts])
    w = np.concatenate([p[2] for p in parts])
    order = np.lexsort((lo, hi))
    ehi, elo, ew = cpu.export_edges()
    assert np.array_equal(hi[order], ehi) and np.array_equal(lo[order], elo) and np.array_equal(w[order], ew)


@pytest.mark.parametrize("k,W,rc", [(31, 2, True), (31, 8, True), (23, 3, True), (27, 4, False), (31, 1, True)])
def test_superkmer_exchange_ranks_emulated_on_one_gpu(K, k, W, rc):
    """super-k-mer exchange: records into the owners' receive buckets, unrolled there into the staged
    sub-table buckets; a uniform batch, a ragged one with rejected reads, and a low-complexity one"""
    from oracle import oracle as O
    seed, G, L, n = 4321 + k, 120_000, 150, 3000 * W
    host = O.synth_reads(seed, G, L, 5000, 0, n)
    rng = np.random.default_rng(k * 10 + W)
    ragged = H.random_reads(rng, 400 * W, k, k + 200, n_rate=0.1) + ["ACGT" * 40, "AT" * 70, "A" * (k + 3)] * W
    ragged = ragged[: len(ragged) // W * W]
    low = np.frombuffer((b"A" * L) * (500 * W), dtype=np.uint8)
    cpu = O.OracleGIR(k)
    cpu.add_reads(host, np.arange(n + 1, dtype=np.uint64) * L, rc)
    cpu.add_reads(*H.batch_of(ragged), rc)
    cpu.add_reads(low, np.arange(500 * W + 1, dtype=np.uint64) * L, rc)
    s = torch.cuda.current_stream().cuda_stream
    gs = [K.GpuGIR(k, rc, world_size=W, rank=r, force_pages=True, stream=s) for r in range(W)]
    sent = 0
    for data, per in ((host, n // W), (low, 500)):
        batches = []
        for r in range(W):
            d = torch.from_numpy(data[r * per * L:(r + 1) * per * L].copy()).cuda()
            offs = torch.arange(0, (per + 1) * L, L, dtype=torch.int64, device="cuda")
            batches.append((d, offs, per, per * L))
        _, keys = _skm_exchange_emulated(K, gs, batches, k)
        sent += sum(keys)
    assert sent == (n + 500 * W) * (L - k + 1)  # the senders' key counts are exact
    per = len(ragged) // W
    batches = []
    for r in range(W):
        b, o = H.batch_of(ragged[r * per:(r + 1) * per])
        batches.append((torch.from_numpy(b).cuda(), torch.from_numpy(o.astype(np.int64)).cuda(), per, int(o[-1])))
    _skm_exchange_emulated(K, gs, batches, k)
    _assert_shards_equal_oracle(gs, cpu, lambda g, h, x: g.mg_skm_owner_of(h, x))
    assert sum(g.info()["page_updates"] for g in gs) > 0
    for g in gs:
        g.close()


def test_superkmer_exchange_spill_route(K):
    """receive buckets sized for a tenth of the batch: most records take the spill route (grouped by
    owner, delivered as flat arrays); growth without a hint on the way"""
    from oracle import oracle as O
    k, W, L = 31, 3, 100
    n = 6000 * W
    host = O.synth_reads(99, 80_000, L, 10000, 0, n)
    cpu = O.OracleGIR(k)
    cpu.add_reads(host, np.arange(n + 1, dtype=np.uint64) * L, True)
    s = torch.cuda.current_stream().cuda_stream
    gs = [K.GpuGIR(k, True, world_size=W, rank=r, force_partition=True, stream=s) for r in range(W)]
    per = n // W
    batches = []
    for r in range(W):
        d = torch.from_numpy(host[r * per * L:(r + 1) * per * L].copy()).cuda()
        offs = torch.arange(0, (per + 1) * L, L, dtype=torch.int64, device="cuda")
        batches.append((d, offs, per, per * L))
    spilled, _ = _skm_exchange_emulated(K, gs, batches, k, claim=1000)
    assert spilled > 0
    _assert_shards_equal_oracle(gs, cpu, lambda g, h, x: g.mg_skm_owner_of(h, x))
    for g in gs:
        g.close()


@pytest.mark.parametrize("k,shards,rc", [(31, 2, True), (31, 4, True), (27, 8, True), (40, 3, False), (63, 2, True),
                                          (32, 4, False), (23, 5, False)])
def test_one_handle_over_several_shards(K, k, shards, rc, tmp_path):
    """ktg_config.n_devices: ONE handle whose table is hash-sharded over `shards` devices (here all on GPU 0),
    fed through the same ktg_add_reads / ktg_create_from_files, answering every query for the whole graph.
    Direct or key exchange below 4 shards and for k outside 23..31, super-k-mer records otherwise.  Against the oracle:
    counters, digest, sorted edges, node / degree statistics, the exported graph (identical to a one-GPU
    handle's, node numbering included), filter, standardize, reset, several calls, file input, the short read."""
    rng = np.random.default_rng(100 * k + shards)
    genome = "".join(rng.choice(list("ACGT"), size=50_000))
    seqs = H.random_reads(rng, 9000, max(k, 50), 170, genome=genome, n_rate=0.02)
    seqs += ["T" * (k + 30), "A" * (k + 9), "AT" * 70, "ACGT" * 40] * 3
    cpu = _oracle(seqs, k, rc)
    bases, offsets = H.batch_of(seqs)
    ids = [0] * shards
    g = K.GpuGIR(k, rc, device_ids=ids, options={"chunk_mb": 1})  # several chunks per call
    assert g.add_reads(bases, offsets) == (cpu.accepted_reads, cpu.accepted_bytes)
    _assert_same(g, cpu)
    # the direct exchange (the sender does the owner's level-1 partition too: the default below 4 shards where
    # the geometries agree) forced on and off
    for direct in (1, 0):
        g2 = K.GpuGIR(k, rc, device_ids=ids, options={"chunk_mb": 1, "mg_direct": direct})
        assert g2.add_reads(bases, offsets) == (cpu.accepted_reads, cpu.accepted_bytes)
        assert g2.digest() == cpu.digest() and g2.counts() == cpu.counts(), direct
        g2.close()
    one = K.GpuGIR(k, rc)
    one.add_reads(bases, offsets)
    ga, gb = g.export_graph(), one.export_graph()
    assert sorted(ga) == sorted(gb)
    for name in ga:
        assert np.array_equal(ga[name], gb[name]), name
    assert g.info()["windows_inserted"] == one.info()["windows_inserted"]
    g.remove_weak_edges(2), cpu.remove_weak_edges(2)
    _assert_same(g, cpu)
    g.standardize_edges(40 * len(genome), k, 3), cpu.standardize_edges(40 * len(genome), k, 3)
    _assert_same(g, cpu, full_stats=False)
    with pytest.raises(K.KatomeError):
        g.standardize_edges(k - 1, k, 1)  # G < k: degenerate (standardizer.rs:123-127)
    # reset, then the same reads in three calls
    g.reset()
    cpu = _oracle(seqs, k, rc)
    third = len(seqs) // 3
    for part in (seqs[:third], seqs[third:2 * third], seqs[2 * third:]):
        g.add_reads(*H.batch_of(part))
    _assert_same(g, cpu, full_stats=False)
    g.close(), one.close()
    # Build::create from files on a sharded handle (builder.rs:42-54)
    fq = H.write_fastq(tmp_path / "r.fastq", seqs)
    g, nbytes = K.GpuGIR.create([fq], "fastq", rc, 0, k=k, device_ids=ids)
    assert nbytes == cpu.accepted_bytes
    _assert_same(g, cpu, full_stats=False)
    g.close()
    # an accepted read shorter than k voids the whole build (hm_gir.rs:40), whichever shard meets it
    g = K.GpuGIR(k, rc, device_ids=ids)
    with pytest.raises(K.ReadTooShort):
        g.add_reads(*H.batch_of(seqs[:500] + ["ACGTA"] + seqs[500:900]))
        g.finalize()
    with pytest.raises(K.KatomeError):
        g.digest()
    g.reset()
    g.add_reads(*H.batch_of(seqs[:100]))
    assert g.digest() == _oracle(seqs[:100], k, rc).digest()
    g.close()
    # per-shard primitives are not offered on such a handle; bad device lists fail loudly
    g = K.GpuGIR(k, rc, device_ids=ids)
    with pytest.raises(K.KatomeError):
        g.mg_plan(1000)
    g.close()
    with pytest.raises(K.KatomeError):
        K.GpuGIR(k, rc, device_ids=[0, 99])


def test_sharded_handle_with_skewed_input_takes_the_spill_route(K):
    """poly-A reads send most windows to one owner: its receive bucket overflows and the rest travels by the
    slow route (grouped by owner, through the host); nothing is lost, in either exchange"""
    from oracle import oracle as O
    rng = np.random.default_rng(77)
    genome = "".join(rng.choice(list("ACGT"), size=30_000))
    seqs = [genome[i:i + 120] for i in rng.integers(0, len(genome) - 120, size=20_000)]
    for i in range(0, len(seqs), 2):
        seqs[i] = "A" * 120 if i % 4 else "ACACACAC" * 15
    bases, offsets = H.batch_of(seqs)
    for k, shards in ((31, 2), (31, 4), (40, 4)):
        m = O.MtCounter(k, True)
        m.add_reads(bases, offsets)
        g = K.GpuGIR(k, True, device_ids=[0] * shards)
        assert g.add_reads(bases, offsets) == m.counters()
        assert g.digest() == m.digest(), (k, shards)
        g.close()


def test_host_mirror_of_owner_matches_device(K):
    from katome_b200 import hashing
    rng = np.random.default_rng(9)
    for k, W in ((31, 8), (40, 3), (64, 5)):
        g = K.GpuGIR(k, True, world_size=W, rank=0)
        lo = rng.integers(0, 2**63, 300, dtype=np.uint64)
        hi = rng.integers(0, 2**(2 * k - 64) if k > 32 else 1, 300, dtype=np.uint64) if k > 32 else np.zeros(300, np.uint64)
        if k <= 32:
            lo &= np.uint64((1 << (2 * k)) - 1)
        want = hashing.owner_of(hi, lo, k, W, True)
        got = [g.owner_of(int(h), int(l)) for h, l in zip(hi.tolist(), lo.tolist())]
        assert got == want.tolist()
        g.close()


def test_full_size_properties(K):
    """Size-independent properties on a run too large for the oracle: weight conservation,
    batch-split invariance, filter idempotence and monotonicity (BASELINE config 2 shape)."""
    from katome_b200.workloads import Workload
    wl = Workload("mini-C2", 1, 1_000_000, 100, 20, 5000, 31)
    n, L = wl.n_reads, wl.read_len
    d = torch.empty(n * L, dtype=torch.uint8, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    K.synth_reads_device(d, wl.seed, wl.genome_len, L, wl.err_ppm, 0, n, stream=s)
    offs = torch.arange(0, (n + 1) * L, L, dtype=torch.int64, device="cuda")
    g = K.GpuGIR(31, True, stream=s, edges_count=wl.expected_distinct_edges())
    assert g.add_reads_device(d, offs, n, n * L, want_counts=True) == (n, n * L)
    D, E, S, M = g.digest()
    assert S == 2 * wl.n_windows  # every window adds 1 to each strand (2 to a palindrome)
    assert 0.8 * wl.expected_distinct_edges() < E < 1.2 * wl.expected_distinct_edges()
    # same reads in two batches, no hint, partitioned path
    g2 = K.GpuGIR(31, True, stream=s, force_partition=True)
    h = n // 2
    g2.add_reads_device(d, offs, h, h * L)
    g2.add_reads_device(d[h * L:], offs[:n - h + 1], n - h, (n - h) * L)
    assert g2.digest() == (D, E, S, M)
    g.remove_weak_edges(3)
    d3 = g.digest()
    assert d3[1] < E and d3[2] < S
    g.remove_weak_edges(3)
    assert g.digest() == d3
    g2.remove_weak_edges(2)
    g2.remove_weak_edges(3)
    assert g2.digest() == d3


def test_baseline_config2_full_size(K):
    """BASELINE config 2 at its full size (4.6 Mbp, 100 bp, 100x, 0.5 % errors, k=31: 322 M windows),
    far beyond what the oracle can run: size-independent properties only."""
    from katome_b200.workloads import C2 as wl
    n, L = wl.n_reads, wl.read_len
    d = torch.empty(n * L + 64, dtype=torch.uint8, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    K.synth_reads_device(d, wl.seed, wl.genome_len, L, wl.err_ppm, 0, n, stream=s)
    offs = torch.arange(0, (n + 1) * L, L, dtype=torch.int64, device="cuda")
    g = K.GpuGIR(wl.k, True, stream=s)  # no capacity hint: sized from the sketch
    assert g.add_reads_device(d, offs, n, n * L, want_counts=True) == (n, n * L)
    D, E, S, M = g.digest()
    assert S == 2 * wl.n_windows                      # every window adds 1 to each strand
    assert 0.8 * wl.expected_distinct_edges() < E < 1.1 * wl.expected_distinct_edges()  # 97.7 M of ~109 M predicted
    assert g.info()["page_updates"] >= 1
    # the same reads in three uneven batches through L2 atomics only: identical table
    g2 = K.GpuGIR(wl.k, True, stream=s, no_pages=True, edges_count=wl.expected_distinct_edges())
    cuts = [0, n // 7, n // 2, n]
    for a, b in zip(cuts[:-1], cuts[1:]):
        g2.add_reads_device(d[a * L:], offs[: b - a + 1], b - a, (b - a) * L)
    assert g2.digest() == (D, E, S, M) and g2.info()["page_updates"] == 0
    # reverse_complement = false: the windows of one strand only
    g3 = K.GpuGIR(wl.k, False, stream=s)
    g3.add_reads_device(d, offs, n, n * L)
    assert g3.digest()[2] == wl.n_windows
    g3.close()
    # filter: idempotent, monotone, and what survives is what the weights say
    g.remove_weak_edges(5)
    d5 = g.digest()
    g2.remove_weak_edges(3)
    g2.remove_weak_edges(5)
    assert g2.digest() == d5 and d5[1] < E and d5[2] < S
    g.remove_weak_edges(5)
    assert g.digest() == d5
    nodes, edges = g.counts()
    assert edges == d5[1] and 0 < nodes <= edges + 2 * 4_600_000
    g.close(), g2.close()


def _gpu_and_oracle_stages(K, wl, n_reads, rc=True, **kw):
    """The three stages of tests/golden/make_baseline_digests.py on the GPU (resident input, ONE call) and on
    the multi-threaded CPU oracle (oracle/katome_oracle_mt.c, digest-equal to the faithful port:
    tests/test_oracle_golden.py) fed with the very bytes the device generator wrote."""
    import sys
    sys.path.insert(0, H.ROOT)
    from oracle import oracle as O
    L = wl.read_len
    d = torch.empty(n_reads * L + 64, dtype=torch.uint8, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    K.synth_reads_device(d, wl.seed, wl.genome_len, L, wl.err_ppm, 0, n_reads, stream=s)
    offs = torch.arange(0, (n_reads + 1) * L, L, dtype=torch.int64, device="cuda")
    g = K.GpuGIR(wl.k, rc, stream=s, **kw)
    assert g.add_reads_device(d, offs, n_reads, n_reads * L, want_counts=True) == (n_reads, n_reads * L)
    m = O.MtCounter(wl.k, rc)
    step = (256 << 20) // L
    for r0 in range(0, n_reads, step):
        r1 = min(n_reads, r0 + step)
        m.add_reads(d[r0 * L: r1 * L].cpu().numpy(), np.arange(r1 - r0 + 1, dtype=np.uint64) * L)
    assert m.counters() == (n_reads, n_reads * L)
    gpu, cpu = {}, {}
    gpu["built"], cpu["built"] = list(g.digest()), list(m.digest())
    g.remove_weak_edges(2), m.remove_weak_edges(2)
    gpu["filtered"], cpu["filtered"] = list(g.digest()), list(m.digest())
    g.standardize_edges(64 * wl.genome_len, wl.k, 3), m.standardize_edges(64 * wl.genome_len, wl.k, 3)
    gpu["standardized"], cpu["standardized"] = list(g.digest()), list(m.digest())
    info = g.info()
    g.close()
    return gpu, cpu, info


def test_baseline_config2_against_the_oracle_at_full_size(K):
    """ALL of BASELINE config 2 (322 M windows, k=31, reverse_complement=true -- the bench workload): the GPU
    digest equals the CPU oracle's after the build, after remove_weak_edges(2) and after
    standardize_edges(64 G, k, 3), and both equal the committed golden digests."""
    import json, os
    from katome_b200.workloads import C2 as wl
    gpu, cpu, info = _gpu_and_oracle_stages(K, wl, wl.n_reads)
    assert gpu == cpu
    gold = json.load(open(os.path.join(H.ROOT, "tests", "golden", "baseline_digests.json")))["workloads"]["c2"]
    for stage in ("built", "filtered", "standardized"):
        assert gpu[stage] == gold[stage], stage
    assert info["page_updates"] >= 1  # the two-level partition + page sweep is what ran


@pytest.mark.parametrize("name,n_reads", [("c3k63", 2_000_000), ("c3", 1_500_000)])
def test_baseline_config3_slice_against_the_oracle(K, name, n_reads):
    """The first reads of BASELINE config 3 (46 Mbp, 150 bp; k=63 = u128 keys, and k=31): 176 M / 180 M windows
    through the paged path, stage by stage against the CPU oracle."""
    from katome_b200.workloads import BY_NAME
    wl = BY_NAME[name]
    gpu, cpu, info = _gpu_and_oracle_stages(K, wl, n_reads)
    assert gpu == cpu
    assert gpu["built"][2] == 2 * n_reads * wl.windows_per_read and info["page_updates"] >= 1


def _both_fastq_parsers(K, path, k, rc, chunk_kb=None):
    """Build::create through the device-side FASTQ parser and through the host reader (its twin)."""
    out = []
    for host in (False, True):
        opts = {"host_parse": 1} if host else ({"fastq_chunk_kb": chunk_kb} if chunk_kb else {})
        try:
            g, nbytes = K.GpuGIR.create([path] if isinstance(path, str) else path, "fastq", rc, 0, k=k, options=opts)
            out.append(("ok", nbytes, g.digest(), g.counts()))
            g.close()
        except K.KatomeError as e:
            out.append(("error", str(e)))
    return out


def test_device_fastq_parser_matches_host_reader(K, tmp_path):
    """FASTQ records cut on the device == the host reader (rust-bio 0.10 semantics:
    strict 4-line records, '@' header, sequence line trimmed on the right, EOF inside a record fails)"""
    rng = np.random.default_rng(77)
    seqs = H.golden_seqs("data2")  # 125 reads, 33 with an N
    ragged = H.random_reads(rng, 400, 45, 160, n_rate=0.05)

    def fastq(reads, nl="\n", final_nl=True, pad=""):
        recs = []
        for i, sq in enumerate(reads):
            qual = "@" * len(sq) if i % 3 == 0 else "I" * len(sq)  # a quality line may start with '@'
            recs.append(f"@read{i} some/description{nl}{sq}{pad}{nl}+{'read%d' % i if i % 2 else ''}{nl}{qual}{nl}")
        text = "".join(recs)
        return text if final_nl else text[: -len(nl)]

    files = {
        "plain": fastq(seqs), "crlf": fastq(seqs, nl="\r\n"), "no_final_newline": fastq(seqs, final_nl=False),
        "trailing_blanks": fastq(seqs, pad=" \t"), "ragged": fastq(ragged), "empty": "",
        "one_record": fastq(seqs[:1]), "many_blocks": fastq(ragged * 12),
    }
    bad = {
        "no_at_first": "read0\nACGT\n+\nIIII\n", "no_at_later": fastq(seqs[:5]) + "read5\nACGT\n+\nIIII\n",
        "cut_after_header": fastq(seqs[:4]) + "@x\n", "cut_after_seq": fastq(seqs[:4]) + "@x\nACGT\n",
        "cut_after_plus": fastq(seqs[:4]) + "@x\nACGT\n+\n", "blank_line_at_end": fastq(seqs[:4]) + "\n",
        "garbage_tail": fastq(seqs[:4]) + "xyz",
    }
    for name, text in {**files, **bad}.items():
        p = tmp_path / f"{name}.fastq"
        p.write_bytes(text.encode())
        # tiny chunks: records carried across chunk boundaries; 64 KiB: the threaded block reader with a carry
        for chunk_kb in (None, 1, 3, 64):
            dev, host = _both_fastq_parsers(K, str(p), 40, True, chunk_kb)
            assert dev == host, (name, chunk_kb, dev, host)
            assert (dev[0] == "error") == (name in bad), (name, dev)
    # against the oracle, and several files in one build
    cpu = _oracle(seqs + ragged, 40, True)
    p1, p2 = tmp_path / "plain.fastq", tmp_path / "ragged.fastq"
    dev, host = _both_fastq_parsers(K, [str(p1), str(p2)], 40, True, 2)
    assert dev == host and dev[2] == cpu.digest() and dev[1] == cpu.accepted_bytes and dev[3] == cpu.counts()


def test_device_fasta_parser_matches_host_reader(K, tmp_path):
    """create_fasta (builder.rs:118-140) with the records cut on the device == the host reader (rust-bio 0.10
    semantics: '>' opens a record, the other lines are appended after trimming trailing whitespace, the first
    line must be a header): wrapped and unwrapped sequences, CRLF, blank lines, no final newline, a record
    that is larger than a chunk (the chunk grows), empty files, errors."""
    rng = np.random.default_rng(78)
    seqs = H.golden_seqs("data2")
    ragged = H.random_reads(rng, 300, 45, 400, n_rate=0.05)

    def fasta(reads, width=60, nl="\n", final_nl=True, pad=""):
        out = []
        for i, sq in enumerate(reads):
            out.append(f">read{i} some description{nl}")
            w = width if width else max(1, len(sq))
            out.extend(sq[j:j + w] + pad + nl for j in range(0, len(sq), w))
        text = "".join(out)
        return text if final_nl else text[: -len(nl)]

    long_read = "".join(rng.choice(list("ACGT"), size=9000))
    files = {
        "wrapped": fasta(seqs), "one_line": fasta(seqs, width=0), "crlf": fasta(seqs, nl="\r\n"),
        "no_final_newline": fasta(seqs, final_nl=False), "trailing_blanks": fasta(seqs, pad=" \t"),
        "ragged": fasta(ragged, width=70), "empty": "", "one_record": fasta(seqs[:1]),
        "blank_lines_inside": fasta(seqs[:20]).replace("\nA", "\n\nA", 5),
        "larger_than_a_chunk": fasta([long_read] + seqs[:10] + [long_read[:5000]], width=80),
    }
    bad = {
        "no_header_first": "ACGT\n>r\nACGT\n", "blank_first": "\n" + fasta(seqs[:3]),
        "header_only_at_end": fasta(seqs[:4]) + ">lonely\n",  # an empty read: "Read is too short!" (hm_gir.rs:40)
    }

    def both(path, chunk_kb):
        out = []
        for opts in ({"fastq_chunk_kb": chunk_kb} if chunk_kb else {}, {"host_parse": 1}):
            try:
                g, nbytes = K.GpuGIR.create([path] if isinstance(path, str) else path, "fasta", True, 0, k=40, options=opts)
                out.append(("ok", nbytes, g.digest(), g.counts()))
                g.close()
            except K.KatomeError as e:
                out.append(("error", e.code))
        return out

    for name, text in {**files, **bad}.items():
        p = tmp_path / f"{name}.fasta"
        p.write_bytes(text.encode())
        for chunk_kb in (None, 1, 3):
            dev, host = both(str(p), chunk_kb)
            assert dev == host, (name, chunk_kb, dev, host)
            assert (dev[0] == "error") == (name in bad), (name, dev)
    cpu = _oracle(seqs + ragged, 40, True)
    dev, host = both([str(tmp_path / "wrapped.fasta"), str(tmp_path / "ragged.fasta")], 2)
    assert dev == host and dev[2] == cpu.digest() and dev[1] == cpu.accepted_bytes and dev[3] == cpu.counts()


# ------------------------------------------------------------------ BFCounter input
@pytest.mark.parametrize("k,rc,t", [(31, True, 0), (40, False, 3), (6, True, 2), (63, True, 5), (32, False, 0)])
def test_bfcounter_input(K, tmp_path, k, rc, t):
    """create_bfc / add_read_bfc on the GPU table == the oracle's restatement: weights, threshold
    pre-filter, accepted bytes, a palindrome, the all-T k-mer at full key width; then reads on top"""
    from oracle import oracle as O
    from tests.test_oracle_golden import _bfc_lines
    rng = np.random.default_rng(100 + k)
    lines = _bfc_lines(rng, k, 2000)
    if k % 2 == 0 and rc:
        lines.append(("ACG" * (k // 6) + "CGT" * (k // 6), 7))
    if not rc:
        lines.append(("T" * k, 9))
    path = tmp_path / "bfc.txt"
    path.write_text("".join(f"{s}\t{w}\n" for s, w in lines))
    cpu, total = O.OracleGIR.create_bfc(k, [str(path)], rc, t)
    g, gtotal = K.GpuGIR.create([str(path)], "bfcounter", rc, t, k=k)
    assert gtotal == total
    _assert_same(g, cpu)
    # the batch entry point, with the threshold applied on the device, and reads mixed in
    g2 = K.GpuGIR(k, rc)
    kmers = np.frombuffer("".join(s for s, _ in lines).encode(), dtype=np.uint8)
    weights = np.array([w for _, w in lines], dtype=np.uint32)
    nk, nb = g2.add_weighted_kmers(kmers, weights, t)
    assert (nk, nb) == (sum(w >= t for _, w in lines), total)
    _assert_same(g2, cpu)
    reads = H.random_reads(rng, 50, k, k + 60)
    g2.add_reads(*H.batch_of(reads))
    cpu.add_reads(*H.batch_of(reads), rc)
    g2.add_read_bfc(lines[0][0].encode(), 5)
    cpu.add_read_bfc(lines[0][0].encode(), 5, rc)
    _assert_same(g2, cpu)
    g.close()
    g2.close()


def test_bfcounter_input_errors(K, tmp_path):
    for body, code in (("ACGT\t3\n", K._lib.KTG_ERR_SHORT_READ), ("A" * 31 + "\n", K._lib.KTG_ERR_BAD_RECORD),
                       ("A" * 31 + "\tx\n", K._lib.KTG_ERR_BAD_RECORD), ("A" * 30 + "N\t2\n", K._lib.KTG_ERR_BAD_RECORD),
                       ("A" * 33 + "\t2\n", K._lib.KTG_ERR_BAD_RECORD)):
        p = tmp_path / "bad.txt"
        p.write_text(body)
        with pytest.raises(K.KatomeError) as e:
            K.GpuGIR.create([str(p)], "bfcounter", True, 0, k=31)
        assert e.value.code == code, body
    with pytest.raises(K.KatomeError) as e:
        K.GpuGIR.create([str(tmp_path / "missing.txt")], "bfcounter", True, 0, k=31)
    assert e.value.code == K._lib.KTG_ERR_IO
