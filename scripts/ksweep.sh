for k in persistent phased; do TM_KRYLOV=$k timeout 250 python scripts/ls89_tight.py 2>&1 | tail -4; TM_KRYLOV=$k timeout 250 python scripts/ls89_tight.py t106_white 2>&1 | tail -4; done
