set -e
bash tools/dev/ncu_pass.sh
bash tools/dev/ncu_full.sh
