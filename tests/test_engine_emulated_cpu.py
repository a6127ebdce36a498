"""The WHOLE library on the CPU: its host logic (capi.cu, engine.cu: chain planning, workspace layout, certification loops,
statistics reduction, Viterbi path resolution, the C ABI) and its general-N kernels are hostified (tests/emu/hostify.py)
and run on the warp emulator with a fake CUDA runtime (tests/emu/cuda_fake.cpp), driven through the C ABI with numpy
buffers as device memory and compared with the oracle (tests/emu/engine_emu_driver.py).

Mode 0 (the default kernels, which ARE validated on the B200) calibrates the emulation itself.  Modes 1 and 2 select the
opt-in panel family (BHMM_B200_PANEL), which had no GPU time in the round it was written: this is the check of its
integration into the engine -- plan capacities, rows of partial statistics, hand-over certification, the time-chunked
Viterbi with its fallback -- that the kernel-level emulation (test_panel_emulated_cpu.py) cannot give."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
EMU = os.path.join(HERE, 'emu')
CSRC = os.path.join(HERE, '..', 'bhmm_b200', 'csrc')


@pytest.fixture(scope='module')
def engine_emu():
    so = os.path.join(EMU, 'engine_emu.so')
    deps = [os.path.join(EMU, f) for f in ('hostify.py', 'cuda_fake.h', 'cuda_fake.cpp', 'warp_emu.h', 'lane_stubs.cpp',
                                           'build_engine_emu.sh')]
    deps += [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cu', '.h', '.cuh'))]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        if not os.path.exists(os.path.join(os.environ.get('CUDA_HOME', '/usr/local/cuda'), 'include', 'cuda_runtime.h')):
            pytest.skip('CUDA headers not found')
        subprocess.run(['bash', os.path.join(EMU, 'build_engine_emu.sh')], check=True, stdout=subprocess.DEVNULL)
    return so


def _drive(mode, specs, trace=False, tiled_chase=False, exact_scan=False):
    env = dict(os.environ, BHMM_B200_PANEL=str(mode))
    if not exact_scan:
        # the slowly mixing specs are there for the repair-sweep paths; the exact scan (which would replace most of those
        # sweeps, and whose N x 32-thread operator blocks are slow to emulate) has its own small test below
        env['BHMM_B200_EXACT_SCAN'] = '0'
    if tiled_chase:
        env['BHMM_B200_CHASE_TILED'] = '1'              # the link kernel for batches with very many segments, forced
    if trace:
        env['EMU_TRACE'] = '1'
    r = subprocess.run([sys.executable, os.path.join(EMU, 'engine_emu_driver.py')] + specs, env=env, stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert ' ok ' in r.stdout and 'FAIL' not in r.stdout
    return r


def test_default_kernels_on_the_emulator(engine_emu):
    """Calibration: the GPU-validated team kernels, certification and path chase reproduce the oracle on the emulator."""
    r = _drive(0, ['5,40,40', '32,40,40', 's32', '32,40,2,300', 'v10', 'l8'], trace=True)
    assert 'block 32 ' in r.stderr                      # the one-warp team kernels ran


def test_panel_family_through_the_engine(engine_emu):
    """BHMM_B200_PANEL=1: N = 32 on the one-warp panel kernels, N = 21 / 40 on the wide kernels, chunked Viterbi, the Gibbs
    sweep on the family's forward filter, and the time-sharded (C5) layout: owned ranges + halo, statistics add up; a slowly mixing model
    with a two-frame warm-up exercises the failed-certification paths (exact forward fix-ups, backward retries with a longer
    warm-up, Viterbi fix-ups)."""
    r = _drive(1, ['32,40,40', '21,40,40', '40,0,0', 's32', '32,40,2,300', 'l8'], trace=True)
    import re
    # no team chain kernel was launched (one warp per block WITH the transition matrix in dynamic shared memory; the exact
    # scan's k_scan_starts also runs 32-thread blocks, without shared memory: the slowly mixing spec triggers it)
    assert not re.search(r'block 32 smem [1-9]', r.stderr)
    assert 'block 256 ' in r.stderr                     # wide kernels, 8 warps (N = 40)
    # structural ties: the chunked Viterbi flags decisions on its path and the sequential team kernel takes over
    r = _drive(1, ['ties'], trace=True)
    assert 'block 32 ' in r.stderr


def test_wide_kernels_at_n32_and_n100_through_the_engine(engine_emu):
    """BHMM_B200_PANEL=2 (N = 32 on the 4-warp wide kernels) and the C4 state count: 13-warp wide kernels + Viterbi with the
    matrix column in registers."""
    _drive(2, ['32,40,40', '32,40,2,300', 'v32', 'w32'])
    r = _drive(1, ['100,40,40'], trace=True)
    assert 'block 416 ' in r.stderr


def test_exact_scan_on_the_emulator(engine_emu):
    """A slowly mixing 6-state model with a two-frame warm-up: the first repair sweep does not clear the failed hand-overs, the
    transfer-operator scan (scan_kernels.cu) supplies exact starts in both directions, and the E-step still equals the oracle."""
    r = _drive(0, ['6,40,2,300'], trace=True, exact_scan=True)
    assert r.stderr.count('block 192 ') >= 2             # k_chain_operator: 32 N threads, forward and backward


def test_exact_scan_wide_on_the_emulator(engine_emu):
    """The same with 40 states: k_chain_operator_wide / k_scan_starts_wide (two columns per lane, 32 warps taking the operator's
    rows in turn) under the tensor-pipe kernels that serve this N by default (BHMM_B200_PANEL=1; the team kernels pass too,
    EMU_SCAN_TEAM=1 runs them as well)."""
    for mode in ((0, 1) if os.environ.get('EMU_SCAN_TEAM') else (1,)):
        r = _drive(mode, ['e40,40,2,300'], trace=True, exact_scan=True)
        assert r.stderr.count('block 1024 ') >= 2        # k_chain_operator_wide, forward and backward
        assert r.stderr.count('block 64 ') >= 2          # k_scan_starts_wide<2>


def test_mover_on_the_fake_runtime(engine_emu):
    """csrc/transfer.cu compiled against the fake CUDA runtime: pieces that straddle arrays and slots, empty arrays, 1 / 3 / 8
    worker threads, streaming and plain stores, guard bands around every destination.  (tests/test_transfer_cuda.py does the
    same through a real GPU.)"""
    _drive(0, ['mover'])


def test_tiled_chase_link_forced(engine_emu):
    """k_chase_link_tiled (chosen on its own only from 65536 segments, i.e. for C5-sized trajectories) forced for every
    batch: batched Viterbi and Gibbs paths with several ragged trajectories, and a literal Viterbi over 36 segments (two
    tiles), in default and panel mode."""
    _drive(0, ['5,40,40', 'l8x9000'], tiled_chase=True)
    _drive(1, ['32,40,40', 'l8x9000'], tiled_chase=True)
