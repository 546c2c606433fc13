python -m pytest tests/test_gpu_parity.py tests/test_gpu_features.py -x -q -m gpu 2>&1 | tail -2
python scripts/time_step.py 16 3 2 1 0 > /dev/null
for ch in 32 64 128 256 1024; do echo -n "chunk=$ch: "; WPT_MEGA_CHUNK=$ch python scripts/time_step.py 16 4 2 1 0; done
for th in "16 8" "24 12" "28 16"; do set -- $th; echo -n "chunk=64 thi=$1 tlo=$2: "; WPT_MEGA_THI=$1 WPT_MEGA_TLO=$2 python scripts/time_step.py 16 4 2 1 0; done
