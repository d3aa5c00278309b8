"""CPU tests that PIN the oracle (oracle/hercules_oracle.c) to the reference.

1. against the committed golden vectors produced by the unmodified reference (ref_dump):
   tables (eTable, nTable, K1, K2, a/b bases) bit-for-bit, whole runs on octor meshes with
   hanging nodes bit-for-bit on full-field snapshots, station rows to their printed 7 digits;
2. against the files the reference's authors shipped (examples/simple/expected-out): source
   forces and station series;
3. when oracle/_ref/libref_kernels.so exists (i.e. /root/reference was compiled here): every
   per-element function against the reference's own stiffness.c / damping.c on random input.
"""
import numpy as np
import pytest

from conftest import load_golden, params_of, rel_l2

SINGLE = ["graded2_rayleigh_eff", "graded2_rayleigh_conv", "graded2_none_eff", "graded2_mass_eff",
          "graded2_bkt", "graded3_rayleigh_eff", "uniform_rayleigh_eff",
          "test1_homogeneous",      # BASELINE.json configs[0] (examples/test1 values, 500 steps)
          "graded2_bkt_qk",         # BKT with finite Qk: shear AND kappa memory variables active
          "basin_rayleigh_eff",     # laterally varying model: hanging nodes on faces / edges of every orientation
          "basin_corner_rayleigh_eff"]   # ... and on the domain's absorbing faces and edges


def snapshots(g, which="tm1"):
    pre = which + "_step"
    return {int(k[len(pre):]): v for k, v in g.items() if k.startswith(pre)}


@pytest.mark.parametrize("name", SINGLE)
def test_tables_bit_exact(oracle, name):
    g = load_golden(name); P = params_of(g)
    K1, K2 = oracle.compute_K()
    assert np.array_equal(K1.reshape(8, 8, 9), g["K1"]) and np.array_equal(K2.reshape(8, 8, 9), g["K2"])
    a, b = oracle.compute_setab(P["damping"], P["freq"])
    assert a == P["abase"] and b == P["bbase"]
    E, N = g["elem_lnid"].shape[0], g["nTable"].shape[0]
    ed = np.ascontiguousarray(g["elem_edata"]).copy()
    eT, nT = np.zeros((E, 4)), np.zeros((N, 7))
    rc = oracle.lib().ho_solver_init_tables(
        E, N, np.ascontiguousarray(g["elem_lnid"]).reshape(-1), np.ascontiguousarray(g["elem_level"]),
        ed.reshape(-1), np.ascontiguousarray(g["node_ticks"]).reshape(-1),
        np.ascontiguousarray(g["domain_ticks"]), P["dt"], P["dt2"], a, b, P["thr_damping"],
        P["thr_vpvs"], eT.reshape(-1), nT.reshape(-1))
    assert rc == 0
    dn = np.ascontiguousarray(g["dnode"])
    oracle.lib().ho_compute_adjust(dn.shape[0], dn.reshape(-1), nT.reshape(-1), 7, oracle.DISTRIBUTION)
    assert np.array_equal(eT, g["eTable"])
    assert np.array_equal(nT, g["nTable"])


@pytest.mark.parametrize("name", SINGLE)
def test_whole_run_bit_exact(oracle, name):
    g = load_golden(name); P = params_of(g)
    m = oracle.Mesh.from_dump(g)
    snaps = snapshots(g)
    s, got = oracle.run(m, P["steps"], P["damping"], P["stiffness"], P["freq"], P["dt"],
                        g["loaded_lnid"], g["forces"], snapshot_steps=set(snaps))
    assert len(snaps) >= 3
    for k, ref in snaps.items():
        assert np.array_equal(got[k], ref), f"step {k}: rel L2 {rel_l2(got[k], ref):.3e}"
    assert np.abs(snaps[max(snaps)]).max() > 0


def test_accelerations_tm3(oracle):
    g = load_golden("graded2_accel"); P = params_of(g)
    assert P["print_accel"] == 1
    m = oracle.Mesh.from_dump(g)
    snaps = snapshots(g)
    s, got = oracle.run(m, P["steps"], P["damping"], P["stiffness"], P["freq"], P["dt"],
                        g["loaded_lnid"], g["forces"], snapshot_steps=set(snaps), accel=True)
    for k, ref in snaps.items():
        assert np.array_equal(got[k], ref)


def station_series(oracle, g, nsteps):
    P = params_of(g)
    m = oracle.Mesh.from_dump(g)
    st = oracle.State(m, bkt=(P["damping"] == oracle.BKT))
    xi = np.array([[-1, 1, -1, 1, -1, 1, -1, 1], [-1, -1, 1, 1, -1, -1, 1, 1],
                   [-1, -1, -1, -1, 1, 1, 1, 1]], float)
    nodes, loc = g["station_nodes"][:, 1:], g["station_local"]
    phi = np.prod(1 + xi[None, :, :] * loc[:, :, None], axis=1) / 8      # psolve.c:6712-6715
    out = np.zeros((nsteps, nodes.shape[0], 3))
    for k in range(nsteps):
        # top of step k after the swap: tm1 (= st.tm2 before step() swaps) is u(t_k)
        out[k] = np.einsum("sj,sjc->sc", phi, st.tm2[nodes])
        oracle.step(m, st, P["damping"], P["stiffness"], P["freq"], P["dt"], g["loaded_lnid"],
                    g["forces"][k] if g["loaded_lnid"].size else None)
    return out


def test_shipped_goldens_examples_simple(oracle):
    """examples/simple/expected-out: forces from the shipped force_process.0.gz, stations from the
    shipped station.N.bz2 (printed with %e: 7 significant digits)."""
    g = load_golden("shipped_simple")
    assert g["forces_rebuilt_max_rel_err"][0] == 0.0        # reference rebuilt here == shipped file
    n = 1500
    got = station_series(oracle, g, n)
    for i in range(5):
        ref = g[f"station{i}"][:n, 1:4]
        ids = g["station_nodes"][:, 0]
        mine = got[:, list(ids).index(i), :]
        scale = np.abs(ref).max()
        assert scale > 1e-3
        # 7 printed digits: relative to each value, 5e-7; tiny values (Z is O(1e-18) noise) are
        # compared against the series scale
        assert np.all(np.abs(mine - ref) <= 6e-7 * np.abs(ref) + 1e-12 * scale), i


@pytest.mark.parametrize("name", ["graded2_rayleigh_eff", "graded2_bkt"])
def test_station_rows(oracle, name):
    g = load_golden(name); P = params_of(g)
    got = station_series(oracle, g, P["steps"])
    ids = list(g["station_nodes"][:, 0])
    for i in range(3):
        ref = g[f"station{i}"][:, 1:4]
        mine = got[:, ids.index(i), :]
        assert np.all(np.abs(mine - ref) <= 6e-7 * np.abs(ref) + 1e-30)


def test_against_reference_kernels(oracle):
    R = oracle.ref_kernels()
    if R is None:
        pytest.skip("oracle/_ref/libref_kernels.so not built (no /root/reference here)")
    L = oracle.lib()
    g = load_golden("graded3_rayleigh_eff")
    ln = np.ascontiguousarray(g["elem_lnid"]).reshape(-1)
    et = np.ascontiguousarray(g["eTable"]).reshape(-1)
    E, N = g["elem_lnid"].shape[0], g["nTable"].shape[0]
    K1, K2 = oracle.compute_K(); K1, K2 = K1.reshape(-1), K2.reshape(-1)
    rng = np.random.default_rng(7)
    for trial in range(3):
        t1, t2 = rng.standard_normal(3 * N), rng.standard_normal(3 * N)
        if trial == 2:                      # exercise the <= 1e-20 early-outs (quake_util.c:36)
            t1[: 3 * N // 2] = 0; t2[: 3 * N // 2] = 0; t1[5] = 1e-21
        pairs = [
            (lambda f: L.ho_addforce_effective(E, ln, et, t1, f), lambda f: R.refk_addforce_effective(E, N, ln, et, t1, f)),
            (lambda f: L.ho_addforce_conventional(E, ln, et, K1, K2, t1, f), lambda f: R.refk_addforce_conventional(E, N, ln, et, K1, K2, t1, f)),
            (lambda f: L.ho_damping_addforce(E, ln, et, K1, K2, t1, t2, f), lambda f: R.refk_damping_addforce(E, N, ln, et, K1, K2, t1, t2, f)),
        ]
        for mine, ref in pairs:
            a, b = np.zeros(3 * N), np.zeros(3 * N)
            mine(a); ref(b)
            assert np.array_equal(a, b)
        ed = (np.abs(rng.standard_normal((E, 14))) * 0.1).astype(np.float32)
        if trial == 1:
            ed[::3, 4:] = 0                 # elements without attenuation (csum == 0 branches)
        ed = ed.reshape(-1)
        c1 = rng.standard_normal((4, 24 * E)); c2 = c1.copy()
        L.ho_calc_conv(E, ln, ed, t1, t2, c1[0], c1[1], c1[2], c1[3], 2.5, 0.001)
        R.refk_calc_conv(E, N, ln, ed, t1, t2, c2[0], c2[1], c2[2], c2[3], 2.5, 0.001)
        assert np.array_equal(c1, c2)
        a, b = np.zeros(3 * N), np.zeros(3 * N)
        L.ho_constant_Q_addforce(E, ln, et, ed, t1, t2, c1[0], c1[1], c1[2], c1[3], a, 2.5, 0.001)
        R.refk_constant_Q_addforce(E, N, ln, et, ed, t1, t2, c2[0], c2[1], c2[2], c2[3], b, 2.5, 0.001)
        assert np.array_equal(a, b)


def test_effective_equals_conventional(oracle):
    """the two stiffness options are the same operator (to rounding)."""
    g = load_golden("graded2_rayleigh_eff")
    L = oracle.lib()
    ln = np.ascontiguousarray(g["elem_lnid"]).reshape(-1); et = np.ascontiguousarray(g["eTable"]).reshape(-1)
    E, N = g["elem_lnid"].shape[0], g["nTable"].shape[0]
    K1, K2 = oracle.compute_K()
    t1 = np.random.default_rng(1).standard_normal(3 * N)
    a, b = np.zeros(3 * N), np.zeros(3 * N)
    L.ho_addforce_effective(E, ln, et, t1, a)
    L.ho_addforce_conventional(E, ln, et, K1.reshape(-1), K2.reshape(-1), t1, b)
    assert rel_l2(a, b) < 1e-14


@pytest.mark.parametrize("name,world", [("basin_rayleigh_eff_np3", 3), ("basin_rayleigh_eff_np4", 4),
                                        ("graded3_rayleigh_eff_np4", 4), ("graded3_rayleigh_eff_np8", 8),
                                        ("uniform_rayleigh_eff_np3", 3),
                                        ("graded2_bkt_np2", 2), ("basin_bkt_np3", 3)])
def test_multirank_oracle_bit_exact(oracle, name, world):
    """The oracle's per-rank arithmetic plus the four schedule_senddata exchanges of a step (psolve.c:4036-4154,
    4945-5079; contribution = += in the receiver's messenger order, sharing = overwrite), carried out in
    process on the tables every rank of the unmodified reference held, reproduces every rank's tm1 snapshots
    bit for bit -- the multi-rank half of the oracle, and the expectation tests/test_multirank_gpu.py holds the
    CUDA path to."""
    from conftest import rank_view
    from hercules_b200.solver import MsgList
    ho = oracle
    g = load_golden(name)
    V = [rank_view(g, r) for r in range(world)]
    P = params_of(V[0])
    M = [ho.Mesh.from_dump(v) for v in V]
    S = [ho.State(m, bkt=(P["damping"] == ho.BKT)) for m in M]
    ML = [{k: MsgList.from_dump(v[k + "_hdr"], v[k + "_map"]) for k in ("dn_c", "dn_s", "an_c", "an_s")} for v in V]
    L = ho.lib()

    def exchange(snd, rcv, arrs, contribution):
        out = []
        for r in range(world):
            m = ML[r][snd]
            off = np.concatenate([[0], np.cumsum(m.nodes)]).astype(int)
            out.append({int(p): arrs[r][m.mapping[off[i]:off[i + 1]]].copy() for i, p in enumerate(m.peer)})
        for r in range(world):
            m = ML[r][rcv]
            off = 0
            for p, n in zip(m.peer.tolist(), m.nodes.tolist()):
                rows = m.mapping[off:off + n]
                if contribution:
                    arrs[r][rows] += out[p][r]
                else:
                    arrs[r][rows] = out[p][r]
                off += n
    snaps = [{int(k[len("tm1_step"):]): a for k, a in v.items() if k.startswith("tm1_step")} for v in V]
    checked = 0
    with np.errstate(all="ignore"):            # a node harbored without mass holds 0/0 until its owner's value arrives
        for k in range(P["steps"]):
            for r in range(world):
                S[r].tm1, S[r].tm2 = S[r].tm2, S[r].tm1
                if k in snaps[r]:
                    assert np.array_equal(S[r].tm1, snaps[r][k]), (r, k)
                    checked += int(np.abs(snaps[r][k]).max() > 0)
                ll = V[r]["loaded_lnid"]
                if ll.size:
                    L.ho_addforce_s(ll.size, np.ascontiguousarray(ll, np.int32),
                                    np.ascontiguousarray(V[r]["forces"][k]).reshape(-1), P["dt2"], S[r].force.reshape(-1))
                ho.add_forces(M[r], S[r], P["damping"], P["stiffness"], P["freq"], P["dt"])
            exchange("dn_c", "dn_s", [s.force for s in S], True)
            for r in range(world):
                L.ho_compute_adjust(M[r].D, M[r].dnode.reshape(-1), S[r].force.reshape(-1), 3, ho.DISTRIBUTION)
            exchange("an_c", "an_s", [s.force for s in S], True)
            for r in range(world):
                L.ho_compute_displacement(M[r].N, M[r].nTable.reshape(-1), S[r].tm1.reshape(-1), S[r].tm2.reshape(-1), None,
                                          S[r].force.reshape(-1))
            exchange("an_s", "an_c", [s.tm2 for s in S], False)
            for r in range(world):
                L.ho_compute_adjust(M[r].D, M[r].dnode.reshape(-1), S[r].tm2.reshape(-1), 3, ho.ASSIGNMENT)
            exchange("dn_s", "dn_c", [s.tm2 for s in S], False)
    assert checked >= world
