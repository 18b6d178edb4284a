"""CPU-only: the task order of the persistent dense solve (ptam_cg_b200/csrc/ldlt_dag.cuh; the reference's
Cholesky<>(mS).backsub(vE), Bundle.cc:457-458).  The kernel hands tasks to resident CTAs through one atomic ticket and
lets them spin on flags, so it is only deadlock-free if the ticket order is a topological order of the dependency
graph.  Checked here on the REAL host table (ptam_bundle_solve_schedule) with a mirror of the device's ticket decode:
  * every block (i,j) of the lower triangle receives the update of every panel k < j exactly once, in panel order;
  * every flag a task waits for is produced by a task with a smaller ticket, or by the chain at a panel not later
    than the task's own; the chain only waits for tasks of earlier panels."""
import ctypes as C
import math

import pytest

NB, TM = 64, 128


def host_table(product, n, tail_tiles):
    nblk = (n + NB - 1) // NB
    off = (C.c_int32 * (nblk + 2))()
    ks = C.c_int(0)
    rc = product.cdll.ptam_bundle_solve_schedule(n, tail_tiles, C.byref(ks), off, nblk + 2)
    assert rc == nblk
    return ks.value, list(off)[: nblk + 1]


def decode(n, nblk, ks, off, t, k):
    """mirror of the worker loop of k_ldlt_dag: ticket -> (kind, panel, parameters); k = the CTA's running panel"""
    if t < off[ks]:
        return ("RU", ks, ks + 2 + t), k
    while t >= off[k + 1]:
        k += 1
    local = t - off[k]
    r0 = (k + 1) * NB
    nt1 = max(0, ((n - r0 + TM - 1) // TM if r0 < n else 0) - 1)
    n_dg, n_ru = max(0, nblk - k - 2), max(0, nblk - k - 3)
    if local < n_dg:
        return ("D", k, k + 2 + local), k
    if local < n_dg + nt1:
        return ("T", k, (1 + local - n_dg, 1)), k
    if local < n_dg + nt1 + n_ru:
        return ("RU", k + 1, k + 3 + (local - n_dg - nt1)), k
    local -= n_dg + nt1 + n_ru
    bi = 1 + int(math.sqrt(local))
    while (bi - 1) * (bi - 1) > local:
        bi -= 1
    while bi * bi <= local:
        bi += 1
    return ("T", k, (bi, local - (bi - 1) * (bi - 1) + 2)), k


def check(product, n, tail_tiles):
    nblk = (n + NB - 1) // NB
    ks, off = host_table(product, n, tail_tiles)
    assert 0 <= ks < max(1, nblk) and ks != 1
    tasks, k = [], ks
    for t in range(off[nblk]):
        task, k = decode(n, nblk, ks, off, t, k)
        tasks.append(task)
    prod, updates = {}, {}
    for idx, (kind, k, prm) in enumerate(tasks):
        if kind == "RU":      # solves block (i,k), then updates block (i,k+1)
            assert k + 2 <= prm < nblk
            prod[("r", prm, k + 1)] = idx
            prod[("u", prm, k + 1, k + 1)] = idx
            updates.setdefault((prm, k + 1), []).append(k)
        elif kind == "D":     # diagonal block (i,i)
            assert k + 2 <= prm < nblk
            prod[("u", prm, prm, k + 1)] = idx
            updates.setdefault((prm, prm), []).append(k)
        else:                 # rows of blocks I, I+1 x column block J, strictly below the diagonal
            bi, bj = prm
            I, J = k + 1 + 2 * bi, k + 1 + bj
            assert I < nblk and 1 <= bj <= 2 * bi
            two, low = I + 1 < nblk, J < I
            if low:
                prod[("u", I, J, k + 1)] = idx
                updates.setdefault((I, J), []).append(k)
            if two:
                prod[("u", I + 1, J, k + 1)] = idx
                updates.setdefault((I + 1, J), []).append(k)
    for k in range(ks, nblk):     # the chain: F(k), then R(k+1,k) and U(k+1,k+1;k)
        prod[("f", k + 1)] = ("chain", k)
        if k + 1 < nblk:
            prod[("r", k + 1, k + 1)] = ("chain", k)
            updates.setdefault((k + 1, k + 1), []).append(k)
    for i in range(ks, nblk):
        for j in range(ks, i + 1):
            assert updates.get((i, j), []) == list(range(ks, j)), (n, i, j)

    def dep(idx, flag, k):
        if flag[-1] <= ks:
            return                # the state the flags start from
        p = prod.get(flag)
        assert p is not None, (n, flag)
        if isinstance(p, tuple):  # produced by the chain at panel p[1]
            assert p[1] < k if idx is None else p[1] <= tasks[idx][1], (n, flag)
        elif idx is None:         # the chain waits for a worker task: it must belong to an earlier panel
            assert tasks[p][1] < k, (n, flag, tasks[p], k)
        else:
            assert p < idx, (n, flag, tasks[p], tasks[idx])

    for idx, (kind, k, prm) in enumerate(tasks):
        if kind == "RU":
            for f in (("f", k + 1), ("u", prm, k, k), ("r", k + 1, k + 1), ("u", prm, k + 1, k)):
                dep(idx, f, k)
        elif kind == "D":
            dep(idx, ("r", prm, k + 1), k)
            dep(idx, ("u", prm, prm, k), k)
        else:
            bi, bj = prm
            I, J = k + 1 + 2 * bi, k + 1 + bj
            two, low = I + 1 < nblk, J < I
            if not low and not two:
                continue
            dep(idx, ("r", I, k + 1), k)
            dep(idx, ("r", J, k + 1), k)
            if two:
                dep(idx, ("r", I + 1, k + 1), k)
                dep(idx, ("u", I + 1, J, k), k)
            if low:
                dep(idx, ("u", I, J, k), k)
    for k in range(ks, nblk - 1):
        dep(None, ("u", k + 1, k, k), k)
        dep(None, ("u", k + 1, k + 1, k), k)
    return ks, len(tasks)


@pytest.mark.parametrize("n", [2, 50, 64, 66, 128, 130, 192, 194, 256, 258, 294, 322, 500, 702, 1000, 1024, 1500, 2994, 3354])
def test_ticket_order_is_a_topological_order(product, n):
    ks, _ = check(product, n, 10 ** 9)     # everything in the persistent kernel
    assert ks == 0
    check(product, n, 120)                 # the default: large systems start with the per-panel schedule
    check(product, n, 16)


def test_hybrid_start_panel(product):
    assert host_table(product, 294, 120)[0] == 0           # C3: all five panels in the persistent kernel
    ks, off = host_table(product, 2994, 120)               # C4: tails of at most 120 tiles from panel 26 on
    assert ks == 26 and off[47] > off[26] > 0
    rem = 2994 - (ks + 1) * NB
    assert ((rem + TM - 1) // TM) ** 2 <= 120 < ((rem + NB + TM - 1) // TM) ** 2


def test_schedule_entry_point_checks_its_arguments(product):
    ks = C.c_int(0)
    off = (C.c_int32 * 4)()
    assert product.cdll.ptam_bundle_solve_schedule(2994, 120, C.byref(ks), off, 4) < 0      # 48 entries needed
    assert product.cdll.ptam_bundle_solve_schedule(-2, 120, C.byref(ks), off, 4) < 0
    assert product.cdll.ptam_bundle_solve_schedule(128, 120, None, off, 4) < 0
    assert product.cdll.ptam_bundle_solve_schedule(128, 120, C.byref(ks), off, 4) == 2 and ks.value == 0
