#!/usr/bin/env python3
"""Regenerates the fixtures under tests/golden/ (run in the dev container, where
/root/reference exists; the GPU box only ever reads the committed outputs).

Inputs : /root/reference/tests/test_files/data{1,2,3}.txt, data_too_short_read.txt
Outputs: c1_data{1,2,3}.seqs      -- the sequence line of every FASTQ record (incl. the reads
                                     with 'N' that the builder must reject)
         reference_pinned.json    -- (a) numbers asserted by the reference's own tests, with
                                     file:line; (b) known-answer vectors of its codec tests;
                                     (c) values derived with the oracle AFTER it reproduced
                                     (a) and (b) (digests, rc=true counts, k=31/33/63 runs)
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/tests/test_files"

from oracle.oracle import OracleGIR  # noqa: E402


def seqs_of(path):
    lines = open(path).read().split("\n")
    return [lines[i + 1].rstrip() for i in range(0, len(lines) - 1, 4) if lines[i].startswith("@")]


def main():
    pinned = {
        "_about": "see make_golden.py; citations are relative to /root/reference",
        "k": 40,
        "reference_pinned": {
            # tests/build.rs:27-28 (accepted bytes, (nodes, edges)), k=40 rc=false t=0
            "data1": {"bytes": 200, "nodes": 62, "edges": 61,
                      # tests/build.rs:46-60
                      "max_edge_weight": 2, "avg_edge_weight": 2.0, "max_in_degree": 1, "max_out_degree": 1,
                      "avg_out_degree": 0.98, "incoming_vert_count": 1, "outgoing_vert_count": 1},
            "data2": {"bytes": 9200, "nodes": 5704, "edges": 5612,
                      # tests/build.rs:61-74
                      "max_edge_weight": 1, "avg_edge_weight": 1.0, "max_in_degree": 1, "max_out_degree": 1,
                      "avg_out_degree": 0.98, "incoming_vert_count": 92, "outgoing_vert_count": 92},
            "data3": {"bytes": 23300, "nodes": 14446, "edges": 14213,
                      # tests/build.rs:76-89
                      "max_edge_weight": 1, "avg_edge_weight": 1.0, "max_in_degree": 1, "max_out_degree": 1,
                      "avg_out_degree": 0.98, "incoming_vert_count": 233, "outgoing_vert_count": 233},
        },
        # tests/pruner.rs:240-251 (HmGIR data1 t=3, data3 t=1), :272 (PtGraph data2 t=2)
        "reference_pinned_filter": [
            {"file": "data1", "threshold": 3, "nodes": 0, "edges": 0},
            {"file": "data3", "threshold": 1, "nodes": 14446, "edges": 14213},
            {"file": "data2", "threshold": 2, "nodes": 0, "edges": 0},
        ],
        "codec_kat": {
            # src/katome/compress.rs:244-248, 277-281
            "compress_edge": [["AGGTCG", [2, 0b00101011, 0b01100000]],
                              # compress.rs:499-520 (adds_char_to_edge preconditions)
                              ["AGGT", [0, 0b00101011]], ["AGGTCGG", [1, 0b00101011, 0b01101000]],
                              ["AGGTC", [3, 0b00101011, 0b01000000]]],
            # compress.rs:334-346
            "encode_fasta_symbol": {"A": 0, "C": 1, "G": 2, "T": 3, "block_ACGT": 0b00011011},
            # compress.rs:558-586: input [0b00101011, 0b01000000] shifted right by 0..8
            "shift_right": {"input": [0b00101011, 0b01000000],
                            "by": [[0, [0b00101011, 0b01000000]], [1, [0b00010101, 0b10100000]],
                                   [2, [0b00001010, 0b11010000]], [3, [0b00000101, 0b01101000]],
                                   [4, [0b00000010, 0b10110100]], [5, [0b00000001, 0b01011010]],
                                   [6, [0b00000000, 0b10101101]], [7, [0b00000000, 0b01010110]],
                                   [8, [0b00101011, 0b01000000]]]},
            # compress.rs:622-681: (input bytes, remainder, output bytes, input string, output string)
            "reverse_compressed_node": [
                [[0b00101011], 0, [0b00010111], "AGGT", "ACCT"],
                [[0b00101011, 0b10110100], 3, [0b10000100, 0b01011100], "AGGTGTC", "GACACCT"],
                [[0b00101011, 0b10110000], 2, [0b00010001, 0b01110000], "AGGTGT", "ACACCT"],
                [[0b00101011, 0b10000000], 1, [0b01000101, 0b11000000], "AGGTG", "CACCT"],
                [[0b10110100, 0b00001111, 0b01010110], 0, [0b01101010, 0b00001111, 0b11100001],
                 "GTCAAATTCCCG", "CGGGAATTTGAC"],
                [[0b10110100, 0b00001111, 0b01010100], 3, [0b10101000, 0b00111111, 0b10000100],
                 "GTCAAATTCCC", "GGGAATTTGAC"],
                [[0b10110100, 0b00001111, 0b01010000], 2, [0b10100000, 0b11111110, 0b00010000],
                 "GTCAAATTCC", "GGAATTTGAC"],
                [[0b10110100, 0b00001111, 0b01000000], 1, [0b10000011, 0b11111000, 0b01000000],
                 "GTCAAATTC", "GAATTTGAC"],
            ],
        },
        # src/katome/algorithms/standardizer.rs:254-279 and :138-141
        "standardize_kat": {"weights": [8, 8, 8, 16, 16, 9, 9, 1, 4, 4, 8, 8, 8, 8], "G": 17, "k": 3, "t": 3,
                            "expect": [1, 1, 1, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1],
                            "ratio": {"G": 10, "k": 0, "s": 10, "l": 0, "p": 1.0}},
    }

    derived = {}
    for name in ("data1", "data2", "data3"):
        src = os.path.join(REF, name + ".txt")
        seqs = seqs_of(src)
        with open(os.path.join(HERE, f"c1_{name}.seqs"), "w") as f:
            f.write("\n".join(seqs) + "\n")
        for k in (21, 31, 32, 33, 40, 63, 64):
            for rc in (False, True):
                g, nbytes = OracleGIR.create(k, [src], "fastq", rc)
                nodes, edges = g.counts()
                st = g.collection_stats()
                entry = {"bytes": nbytes, "nodes": nodes, "edges": edges, "digest": list(g.digest()),
                         "stats": {kk: st[kk] for kk in ("max_edge_weight", "sum_edge_weight", "max_in_degree",
                                                        "max_out_degree", "incoming_vert_count",
                                                        "outgoing_vert_count")}}
                for t in (2, 3):
                    g2, _ = OracleGIR.create(k, [src], "fastq", rc)
                    g2.remove_weak_edges(t)
                    entry[f"filter_t{t}"] = {"counts": list(g2.counts()), "digest": list(g2.digest())}
                derived[f"{name}/k{k}/rc{int(rc)}"] = entry
        if name in pinned["reference_pinned"]:
            ref = pinned["reference_pinned"][name]
            got = derived[f"{name}/k40/rc0"]
            assert (got["bytes"], got["nodes"], got["edges"]) == (ref["bytes"], ref["nodes"], ref["edges"]), name
    pinned["oracle_derived"] = derived
    # the too-short fixture: 1st read has N's (rejected), 2nd is 7 bp -> the build must abort
    pinned["too_short_seqs"] = seqs_of(os.path.join(REF, "data_too_short_read.txt"))
    with open(os.path.join(HERE, "reference_pinned.json"), "w") as f:
        json.dump(pinned, f, indent=1, sort_keys=True)
    print("wrote", len(derived), "derived entries")


if __name__ == "__main__":
    main()
