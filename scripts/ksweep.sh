export TM_MG_AA=0
timeout 100 python scripts/mg_features.py 4 2>&1 | tail -2
OMEGA=0.6 timeout 100 python scripts/mg_features.py 4 2>&1 | tail -1
NU=6 timeout 100 python scripts/mg_features.py 4 2>&1 | tail -1
TM_MG_COARSEST_SWEEPS=3000 timeout 100 python scripts/mg_features.py 4 2>&1 | tail -1
TM_MG_MAX_LEVELS=5 TM_MG_COARSEST_SWEEPS=3000 timeout 100 python scripts/mg_features.py 4 2>&1 | tail -1
WHICH=outer timeout 100 python scripts/mg_features.py 4 2>&1 | tail -2
