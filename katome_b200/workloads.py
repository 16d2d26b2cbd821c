"""Synthetic workloads of BASELINE.json (`configs`), as parameters of the
counter-based read generator (oracle ko_synth_reads / synth_reads_kernel)."""
from __future__ import annotations

from dataclasses import dataclass

SEED_BASE = 0x6B61746F6D65  # "katome"


@dataclass(frozen=True)
class Workload:
    name: str
    config_index: int
    genome_len: int
    read_len: int
    coverage: int
    err_ppm: int
    k: int

    @property
    def seed(self) -> int:
        return SEED_BASE + self.config_index

    @property
    def n_reads(self) -> int:
        return -(-self.genome_len * self.coverage // self.read_len)

    @property
    def windows_per_read(self) -> int:
        return self.read_len - self.k + 1

    @property
    def n_windows(self) -> int:
        return self.n_reads * self.windows_per_read

    def expected_distinct_edges(self) -> int:
        """Both-strand distinct edges, the way the reference counts them."""
        g = 2 * (self.genome_len - self.k + 1)
        bases = self.n_reads * self.read_len
        e = self.err_ppm / 1e6
        err = 2 * e * bases * self.k * (1 - (self.k - 1) / self.read_len)
        return int(g + err)

    def algorithmic_bytes_per_window(self) -> float:
        """input ASCII per window + key + 4 B weight read + 4 B weight write."""
        key = 8 if self.k <= 32 else 16
        return self.read_len / self.windows_per_read + key + 8


C2 = Workload("C2 4.6Mbp/100bp/100x/0.5%/k31", 1, 4_600_000, 100, 100, 5000, 31)
C3_K31 = Workload("C3 46Mbp/150bp/50x/0.5%/k31", 2, 46_000_000, 150, 50, 5000, 31)
C3_K63 = Workload("C3 46Mbp/150bp/50x/0.5%/k63", 2, 46_000_000, 150, 50, 5000, 63)
C5 = Workload("C5 1Gbp/150bp/30x/0.5%/k31", 4, 1_000_000_000, 150, 30, 5000, 31)

BY_NAME = {"c2": C2, "c3": C3_K31, "c3k63": C3_K63, "c5": C5}
