timeout 600 python -m pytest tests/test_gpu_trainer_sequence.py -q -x 2>&1 | tail -12 | cut -c1-300
timeout 600 python profiles/fsptq_recon_c3.py --iters 200 2>&1 | tail -2 | tee gpurun_out/r02_fsptq_recon_c3.json
