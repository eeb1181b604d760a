"""A-vs-B report over benchmark JSON lines, in the shape of the reference's benchmark report
(/root/reference/src/tcn/benchmark/report.py:80-93: one line per label, ``label: 1.00x (a s) - N.NNx (b s)``,
medians of the per-timestep samples).  The GEOS log miner that feeds the reference report is out of scope
(nothing to mine without GEOS); the inputs here are bench.py / sweep.py records.
"""
from __future__ import annotations

import json
import statistics
from dataclasses import dataclass, field
from typing import Dict, List, Sequence


@dataclass
class Entry:
    label: str
    seconds: List[float]

    @property
    def median(self) -> float:
        return statistics.median(self.seconds)


@dataclass
class BenchmarkReport:
    baseline: str
    candidate: str
    lines: List[str] = field(default_factory=list)

    def __str__(self) -> str:
        head = f"Benchmark: {self.candidate} vs {self.baseline} (baseline = 1.00x)"
        return "\n".join([head, "-" * len(head)] + self.lines)


def compare(baseline_name: str, baseline: Dict[str, Sequence[float]], candidate_name: str,
            candidate: Dict[str, Sequence[float]]) -> BenchmarkReport:
    """Per label: median time of each side and the speed-up of the candidate over the baseline."""
    rep = BenchmarkReport(baseline_name, candidate_name)
    for label in baseline:
        if label not in candidate:
            continue
        a, b = Entry(label, list(baseline[label])), Entry(label, list(candidate[label]))
        rep.lines.append(f"{label}: 1.00x ({a.median:.6f}s) - {a.median / b.median:.2f}x ({b.median:.6f}s)")
    return rep


def from_bench_lines(gpu_line: str, reference_line: str) -> BenchmarkReport:
    """bench.py line vs ``bench.py --impl reference`` line: seconds per grid point x level."""
    g, r = json.loads(gpu_line), json.loads(reference_line)
    label = g["config"]["workload"].split(" (")[0]
    base = {label + " [resident fields]": [1.0 / r["value"]], label + " [host buffers, e2e]": [1.0 / r["value"]]}
    cand = {label + " [resident fields]": [1.0 / g["value"]]}
    if g.get("e2e"):
        cand[label + " [host buffers, e2e]"] = [1.0 / g["e2e"]["value"]]
    return compare(f"CPU {r['cpu_baseline']['kind']} ({r['cpu_baseline']['cores']} threads)", base,
                   f"B200 x{g['n_gpus']} ({g['dtype']})", cand)
