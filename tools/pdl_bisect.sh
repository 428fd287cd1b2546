for m in 31 0 1 2 4 1 31; do
  echo "== FHVAE_PDL=$m"
  FHVAE_PDL=$m timeout 300 python -m pytest tests/test_gpu_model.py -q -x -k "tensor_core_modes" 2>&1 | tail -3
done
