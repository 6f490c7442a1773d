"""Multi-GPU host logic on the CPU: the sharded schedule (exchange planning, in-tile qubit permutation
passes, rank-bit predicates) interpreted by the host test double must reproduce the oracle; plus a
world_size-2 torch.distributed/gloo run where every process holds one shard."""
import os
import subprocess
import sys

import numpy as np
import pytest

import helpers
import gpu_quantum_simulator_b200 as q
from gpu_quantum_simulator_b200 import circuits


@pytest.mark.parametrize("precision", [32, 64])
@pytest.mark.parametrize("n,world,depth,seed", [(19, 2, 6, 1), (20, 4, 5, 2), (21, 8, 4, 3)])
def test_sharded_schedule_reproduces_oracle(n, world, depth, seed, precision):
    circ = circuits.random_layered(n, depth=depth, seed=seed)
    got, rep = helpers.sharded_host_run(q.gates_from_circuit(circ), n, world, precision)
    want = helpers.oracle_run_circuit(circ, n)
    assert rep["swaps"] >= 1 and rep["bad_slots"] == 0 and rep["max_conflict"] == 1
    assert np.max(np.abs(got - want)) < 1e-12


@pytest.mark.parametrize("precision", [32, 64])
@pytest.mark.parametrize("n,world,depth,seed", [(19, 2, 6, 1), (20, 4, 5, 2), (21, 8, 4, 3)])
def test_fused_exchange_schedule_reproduces_oracle(n, world, depth, seed, precision):
    """Exchanges fused into the preceding pass: every rank scatters straight into its peers' shards."""
    circ = circuits.random_layered(n, depth=depth, seed=seed)
    got, rep = helpers.sharded_host_run(q.gates_from_circuit(circ), n, world, precision, fused=True)
    want = helpers.oracle_run_circuit(circ, n)
    assert rep["swaps"] >= 1 and rep["bad_slots"] == 0 and rep["max_conflict"] == 1
    assert np.max(np.abs(got - want)) < 1e-12


def test_sharded_superset_and_qft():
    for circ, n, world in ((circuits.random_superset(19, 150, 9), 19, 4), (circuits.qft(19), 19, 2)):
        want = helpers.oracle_run_circuit(circ, n)
        for fused in (False, True):
            got, rep = helpers.sharded_host_run(q.gates_from_circuit(circ), n, world, 32, swap_min_ops=4, fused=fused)
            assert rep["bad_slots"] == 0
            assert np.max(np.abs(got - want)) < 1e-12


def test_single_round_permutation_passes_synchronise_before_the_scatter():
    """an in-place pass that relocates qubits inside its tile in ONE round has no exchange barrier
    between gather and scatter; its descriptor must carry QSB_PASS_SYNC_SCATTER.  (The host double runs threads one
    after the other and cannot see the race, so the flag is asserted on the plan.)"""
    seen = 0
    for seed in range(6):
        for world, n in ((2, 19), (4, 20), (8, 21)):
            circ = circuits.random_layered(n, depth=5, seed=seed)
            r = helpers.ShardedHostRun(q.gates_from_circuit(circ), n, world, 0, 32)
            for i in range(r.steps):
                p = r.props(i)
                if p is None:
                    continue
                risky = p["rounds"] == 1 and not p["out_of_place"] and p["moved"]
                assert bool(p["sync_scatter"]) == bool(risky), (seed, world, i, p)
                seen += risky
            r.finish()
    assert seen > 0, "no single-round permutation pass in the sample: the test checks nothing"


@pytest.mark.timeout(300)
def test_high_exchange_threshold_terminates():
    """Found by the sharded fuzzer once the doubles ran the library's own plan search: with an exchange threshold the tail
    of a circuit never reaches (14 runnable gates, one of the search's candidates) the NCCL-style schedule exchanged
    qubits for ever -- each exchange's permutation pass consumed nothing.  A second exchange without an op consumed since
    the first is now refused while the pass has anything to run; every threshold must plan, and reproduce the oracle."""
    n, world, prec = 18, 4, 32            # fuzz seed 23600015: the 18-qubit QFT on four ranks
    circ = circuits.qft(n)
    want = helpers.oracle_run_circuit(circ, n)
    for fused in (False, True):
        for smo in (14, 20, 40, 0):
            got, rep = helpers.sharded_host_run(q.gates_from_circuit(circ), n, world, prec, swap_min_ops=smo, fused=fused)
            assert rep["bad_slots"] == 0 and np.max(np.abs(got - want)) < 1e-12, (fused, smo)
            assert rep["swaps"] <= 6, (fused, smo, rep)


def test_exchange_count_on_the_34_qubit_workload():
    circ = circuits.random_layered(34, 20, 12345)
    g = q.gates_from_circuit(circ)
    one = q.plan_dry_run(34, g, world_size=1)
    for world in (2, 4, 8):
        st = q.plan_dry_run(34, g, world_size=world)
        assert 1 <= st["swaps"] <= 8
        assert st["passes"] <= one["passes"] + 6
        local = (1 << 34) // world * 8
        assert st["bytes_exchanged"] == st["swaps"] * (local - local // world)


def test_two_process_gloo_run(tmp_path):
    script = os.path.join(os.path.dirname(__file__), "dist_host_worker.py")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", script, str(tmp_path)],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "max_abs_err" in r.stdout
    err = float(r.stdout.split("max_abs_err=")[1].split()[0])
    assert err < 1e-12
