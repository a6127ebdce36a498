"""Parity of the panel family (panel_kernels.cu: N = 32 one-warp kernels, 17 <= N <= 104 wide kernels, register Viterbi,
time-chunked Viterbi) against the oracle, on a GPU.  The family is the default since round 2; the library reads
BHMM_B200_PANEL once per process, hence one child process per mode (`timeout` bounds a possible hang):

* mode 1 (default): N = 32 on the one-warp kernels, N = 21 / 37 / 64 / 100 on the wide kernels, the C4 shape (N = 100, M = 1000)
* mode 2: N = 32 on the 4-warp wide kernels
* mode 0: the SAME cases on the team kernels, the fallback that BHMM_B200_PANEL=0 selects
"""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*flags):
    r = subprocess.run(['timeout', '600', sys.executable, os.path.join(ROOT, 'tests', 'panel_check.py'), '--quick'] + list(flags),
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    print(r.stdout)
    assert r.returncode == 0, r.stdout[-4000:]
    assert 'parity: 0 failure(s)' in r.stdout


def test_panel_kernels_match_oracle():
    _run()


def test_wide_kernels_at_n32_match_oracle():
    _run('--mode=2')


def test_team_kernels_same_cases():
    _run('--team-parity')
