set -x
for cg in 1 3; do for epi in direct bulk; do for dbg in 0 3 1 2; do
K2_PROF_CG=$cg OFB_K2_EPI=$epi OFB_K2_DBG=$dbg timeout 120 python tools/k2_profile.py c5b8 2>&1 | tail -1
done; done; done
