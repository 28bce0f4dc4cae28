"""profiles/r1_parity_sweep.md from the JSONs written on the GPU box by scripts/gpu_family_sweep.sh."""
import glob, json, os
rows = []
for fa in sorted(glob.glob('gpurun_out/sweep_b200_*.json')):
    fb = fa.replace('sweep_b200_', 'sweep_oracle_')
    if not os.path.exists(fb):
        continue
    a, b = json.load(open(fa)), json.load(open(fb))
    if a["iterations"] < 0:
        continue
    rows.append((a, b))
with open('profiles/r1_parity_sweep.md', 'w') as f:
    f.write("# Round-1 parity sweep: B200 backend vs CPU oracle, both driving the reference's unmodified IP loop (B200 box)\n\n")
    f.write("`scripts/gpu_family_sweep.sh` (tests/driver/ipopt_driver; reference TNLPs from examples/ScalableProblems compiled in place).\n"
            "wi = SYMSOLVER_WRONG_INERTIA returns (inertia-correction retries of the reference, IpPDFullSpaceSolver.cpp:541-591).\n\n")
    f.write("| problem | N | KKT dim | IP iterations GPU / oracle | factorisations GPU / oracle | wi GPU / oracle | final objective (GPU) | rel. diff vs oracle | GPU factor ms | GPU solve ms | oracle factor ms |\n|---|---|---|---|---|---|---|---|---|---|---|\n")
    for a, b in rows:
        f.write("| %s | %d | %d | %d / %d | %d / %d | %d / %d | %.15e | %.1e | %.2f | %.2f | %.0f |\n" % (
            a['problem'], a['N'], a['kkt_dim'], a['iterations'], b['iterations'], a['n_factor'], b['n_factor'],
            a['n_wrong_inertia'], b['n_wrong_inertia'], a['objective'],
            abs(a['objective'] - b['objective']) / max(abs(b['objective']), 1e-300),
            1e3 * a['t_factor_s'] / max(a['n_factor'] - 1, 1), 1e3 * a['t_solve_s'] / max(a['n_solve'], 1),
            1e3 * b['t_factor_s'] / max(b['n_factor'] - 1, 1)))
print(len(rows), "rows")
