export TM_MG_AA=0
timeout 120 python scripts/mg_passages_diag.py 4 relax 2>&1 | tail -1
TM_MG_MAX_LEVELS=2 TM_MG_COARSEST_SWEEPS=3000 timeout 120 python scripts/mg_passages_diag.py 4 mg 2>&1 | tail -1
TM_MG_MAX_LEVELS=2 timeout 120 python scripts/mg_passages_diag.py 4 mg 2>&1 | tail -1
timeout 120 python scripts/mg_passages_diag.py 4 mg 2>&1 | tail -1
PICARD=6 timeout 200 python scripts/mg_passages_diag.py 4 mg 2>&1 | tail -2
PICARD=6 TM_MG_MAX_LEVELS=2 TM_MG_COARSEST_SWEEPS=3000 timeout 200 python scripts/mg_passages_diag.py 4 mg 2>&1 | tail -1
