# usage: bash tools/gpu_tests.sh TAG [pytest args...]
TAG=$1; shift
D=gpurun_out/$TAG; mkdir -p $D
python -m pytest "$@" -m gpu -q --durations=8 > $D/tests.log 2>&1; echo "pytest exit $?" >> $D/tests.log
