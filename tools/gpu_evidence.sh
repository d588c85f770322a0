#!/bin/bash
# Evidence run on the GPU box (one B200): compute-sanitizer over the parity tests, the ncu launch list of one
# bench step, and one `ncu --set full` capture each of the contraction (K1) and the rescoring (K2) kernel.
#   gpurun --timeout 1500 -- 'bash tools/gpu_evidence.sh r2'
# Everything lands in gpurun_out/; tools/ncu_summary.py turns the reports into the text files of profiles/.
# Numbers printed by a run under a profiler or sanitizer are never bench values.
set -u
tag=${1:-r2}
what=${2:-all}
out=gpurun_out
mkdir -p $out/sanitizer
export PYTHONUNBUFFERED=1

san() {  # san <tool> <limit seconds> <pytest -k expression>
    local tool=$1 limit=$2 expr=$3 log=$out/sanitizer/${tag}_$1.log
    echo "# compute-sanitizer --tool $tool; pytest -m gpu -k \"$expr\"" > $log
    timeout $limit compute-sanitizer --tool $tool --target-processes all --error-exitcode 99 --print-limit 20 \
        python -m pytest tests/test_gpu_parity.py tests/test_gpu_multirank.py tests/test_pgvector_io.py -m gpu -x -q \
        -p no:cacheprovider -o addopts="" --timeout 0 -k "$expr" >> $log 2>&1
    echo "# exit code $? (99 = the sanitizer reported errors, 124 = time limit)" >> $log
    tail -4 $log
}

if [ "$what" = sanitizer ]; then  # (compute-sanitizer is closed on this pool: exit code 86; kept for pools that have it)
    # memcheck: every kernel family at test size -- the three list widths (<4,.> <8,.> <16,.>), A resident (D <= 512)
    # and streamed (D = 768, 1024), CTA pairs, the slab pipeline with host pieces, the exact scan stages, ingest,
    # half-precision rows, COPY decode, the sharded passes
    san memcheck 900 "fused_path_matches_oracle or cta_pair_path_matches_oracle or two_stage_exact_scan or (slab_pipeline_matches_oracle and pinned and all) or 512_entry_lists or widest_lists or term_bitsets_golden or half_precision_rows or image_term_sets or copy_decode or exact_scan_matches_oracle or first_slab_by_column_groups or (column_sharded and 300-2003) or (fully_sharded and 300-2003) or small_corpus_alignment_records"
    # racecheck (shared-memory hazards) and synccheck (barrier misuse): the mbarrier / TMEM kernels and the block sorts
    san racecheck 600 "(fused_path_matches_oracle and (130-2000-64 or 256-4096-768)) or cta_pair_raw_scores or (exact_scan_matches_oracle and 130-300) or (two_stage_exact_scan and 2500) or 512_entry_lists"
    san synccheck 600 "(fused_path_matches_oracle and (130-2000-64 or 256-4096-768)) or cta_pair_raw_scores or (exact_scan_matches_oracle and 130-300) or (two_stage_exact_scan and 2500) or 512_entry_lists"
    san initcheck 600 "(fused_path_matches_oracle and (1000-5000-512 or 200-3000-1024)) or (slab_pipeline_matches_oracle and device and all) or term_bitsets_golden"
fi

if [ "$what" = all ] || [ "$what" = checked ]; then
    # the library's own checked build (csrc/Makefile `make check`) under the whole GPU suite; the report of
    # tests/conftest.py::pytest_sessionfinish lands in gpurun_out/checked_build_report.txt
    MMALIGN_LIB=$PWD/multimodal-alignment-of-noisy-image-text-pairs-using-weak-supervision_b200/csrc/libmmalign_check.so \
        python -m pytest tests -m gpu -q -p no:cacheprovider > $out/${tag}_checked_pytest.log 2>&1
    tail -4 $out/${tag}_checked_pytest.log
fi

if [ "$what" = all ] || [ "$what" = ncu ]; then
    cmd="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-verify"
    $cmd > $out/${tag}_plain.json 2> $out/${tag}_plain.err || { echo "bench failed without ncu"; tail -5 $out/${tag}_plain.err; exit 1; }
    ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^(mma::|fused_score|select_kernel|gather_kernel|rank_kernel|rescore_kernel|exact_|prep_rows|page_range|metrics_|max_float|iota_|Device)' \
        -c 400 --csv --log-file $out/${tag}_launches.csv $cmd > $out/${tag}_ncu_l.log 2>&1
    # one full capture of each: the timed step's first launch of K1 and of K2 (launches 0 and 1 belong to the warm-up step: whole waves, then the remainder)
    ncu --set full --clock-control none --import-source on -k regex:fused_score_topk -s 2 -c 1 -f \
        -o $out/${tag}_k1 $cmd > $out/${tag}_ncu_k1.log 2>&1
    # K2: the three kernels of the timed step's first slab (the warm-up step launched each twice)
    ncu --set full --clock-control none --import-source on -k regex:'^(select_kernel|gather_kernel|rank_kernel)$' -s 6 -c 3 -f \
        -o $out/${tag}_k2 $cmd > $out/${tag}_ncu_k2.log 2>&1
    ls -la $out/${tag}_k1.ncu-rep $out/${tag}_k2.ncu-rep
fi
