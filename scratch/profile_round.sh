set -e
bash scratch/ncu_pass.sh
bash scratch/ncu_full.sh
