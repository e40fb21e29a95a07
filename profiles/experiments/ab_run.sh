python -m pytest tests -q -m gpu -s -k "f32_vs_f64 or c5_full or merge or strong_scaling_shard or c3_c4_full" 2>&1 | tail -25
