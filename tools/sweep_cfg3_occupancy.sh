# k=16 KNN of the config-3 barycentres for several cell occupancies (SSDR_KNN_OCCUPANCY = points per occupied cell / K)
for o in ${@:-0.3 0.6 0.8 1.0 1.3}; do echo "occupancy $o"; SSDR_KNN_OCCUPANCY=$o python tools/prof_cfg3_knn.py 2>&1 | grep "world 1 rank 0 call 2\|world 8 rank 4 call 2"; done
