# k=16 KNN of the config-3 barycentres for several cell budgets (SSDR_KNN_CELLS_PER_POINT)
for o in ${@:-4 8 16 32}; do echo "cells per point $o"; SSDR_KNN_CELLS_PER_POINT=$o python tools/prof_cfg3_knn.py 2>&1 | grep "world 1 rank 0 call 2\|world 8 rank 4 call 2\|world 8 rank 0 call 2"; done
