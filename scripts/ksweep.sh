echo "== o=8 thick 0.004"; OGRID=8 THICK=0.004 timeout 120 python scripts/mg_passages.py 8 2>&1 | tail -3
echo "== o=8 thick 0.004 AA off"; TM_MG_AA=0 OGRID=8 THICK=0.004 timeout 120 python scripts/mg_passages.py 8 2>&1 | tail -2
echo "== o=4 thick 0.002"; OGRID=4 THICK=0.002 timeout 120 python scripts/mg_passages.py 8 2>&1 | tail -2
echo "== o=8 thick 0.004 factor 16"; OGRID=8 THICK=0.004 timeout 120 python scripts/mg_passages.py 16 2>&1 | tail -2
echo "== o=8 thick 0.004 nu 5"; NU=5 OGRID=8 THICK=0.004 timeout 120 python scripts/mg_passages.py 8 2>&1 | tail -2
