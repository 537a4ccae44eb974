"""How tests/golden/tie_env2.npz was made (needs /root/reference and g++; takes a few minutes).

The fixture holds the EGM points (M, C, V_d in generation order, the last 1000 of 10000) of decision 0 in period it=7
of S1b at BASELINE size -- retirement2 with the shipped parameters at ngridm=10000, ny=100, T=40 -- in two versions:

  ref_*   computed from the reference's own period-8 cell: a flat stretch of eight points with bit-identical C and V
          (next-period consumption is constant there: the constant extrapolation of a run of the period-8 secondary
          envelope, egdst_solver.c:824-827), then a fold-back
  pert_*  the same with the period-8 value at one double point lowered by one ulp (what a different summation order
          of the expectation produces): one point of the flat stretch comes out one ulp lower, the split criterion
          V[i-1] > V[i] (egdst_solver.c:819) fires inside the stretch, and the secondary envelope sees two runs whose
          constant extrapolations coincide -- an exact tie between two functions over a whole interval

Steps: (1) solve S1b with the compiled reference (oracle/ref.py) and keep its cells; (2) import them into the host
emulator build of the kernels (tools/hostemu), perturb V of the double point at M = 6.8279689 in period 8, and re-run
period 7 only (EGDST_SOLVE_FROM=7 EGDST_SOLVE_TO=7) with EGDST_DEBUG_DUMP_IT=7, which writes the per-decision point
lists before the secondary envelope to /tmp/egdst_pt_raw_sd0.bin; (3) keep points 9000..9999.
"""
