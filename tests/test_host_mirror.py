"""Host-side pieces of the reference API that need no GPU: the LANL statistics
table (pdes_lanl_benchmarkV8.js:317-341) and the guard on getOffsetRank."""
import io

import numpy as np

from simian_b200.workloads import LANL_STATE_DTYPE, lanl_report


def test_lanl_report_format_and_totals():
    st = np.zeros(3, dtype=LANL_STATE_DTYPE)
    st["send_count"] = [5, 0, 7]
    st["receive_count"] = [2, 0, 4]
    st["ops_count"] = [200, 0, 410]
    st["ops_min"] = [90.7, np.inf, 80.2]
    st["ops_mean"] = [100.0, 0.0, 102.5]
    st["ops_max"] = [110.9, 0.0, 130.1]
    st["time_sends"][:, :4] = [[1, 2, 1, 1], [0, 0, 0, 0], [3, 2, 1, 1]]
    out = io.StringIO()
    tot = lanl_report(st, 4, out)
    lines = out.getvalue().splitlines()
    # sprintf("%10s %10s %10s %10s %10s %10s    '%s'\n", ...)            :318
    assert lines[0] == " #EntityID     #Sends  #Receives    Ops(Min        Avg       Max)    'Time Bin Sends'"
    # sprintf("%10d %10d %10d %10d %10.5g %10d    %s\n", ...)              :331
    assert lines[1] == "         0          5          2         90        100        110    [1,2,1,1]"
    assert lines[2] == "         1          0          0          0          0          0    [0,0,0,0]"
    assert lines[3] == "         2          7          4         80      102.5        130    [3,2,1,1]"
    assert lines[4].startswith("===================== LANL PDES BENCHMARK  Collected Stats from All Ranks ")
    assert lines[5].split()[:4] == ["#Entities", "#Sends", "#Receives", "OpsCount"]
    assert lines[6].split()[:4] == ["3", "12", "6", "610"]
    assert lines[7] == "=" * 97
    assert tot["sends"] == 12 and tot["receives"] == 6 and tot["ops_count"] == 610
    # gops_mean = (r*m + g*R) / max(1, R + r), folded in entity order          :324
    assert tot["ops_mean"] == (4 * 102.5 + 100.0 * 2) / 6
    assert list(tot["time_sends"]) == [4, 4, 2, 2]
