set -x
nvidia-smi -L
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for args in "tiny16 64 2" "tiny16 1 1" "velodyne64 512 1.5" "velodyne64 2048 2.3 dict(moving=True,dropout=0.05)" "velodyne64 4096 3.1 dict(moving=True)" "kitti64 1024 1.3" "vls128 512 1.3 dict(moving=True,start_firing=40)" "os32_left 256 2.0 dict(moving=True)"; do
  timeout 300 python tests/run_parity.py $args 2>&1 | tail -3
done
