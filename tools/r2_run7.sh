#!/bin/bash
echo "---- non-solve part of the step with 8 groups overlapping (solve skipped: diagnostics, results wrong)"
B2J_DIAG_SKIP_SOLVE=1 tools/r2_solve_ab.sh "0:8 0:1"
tools/profile_round.sh r2 256
