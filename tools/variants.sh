for v in "-DHRT_BQ_WARPS=8 -DHRT_BQ_MIN_CTAS=1 -DHRT_BQ_ALIGN=1" "-DHRT_BQ_WARPS=12 -DHRT_BQ_MIN_CTAS=1 -DHRT_BQ_ALIGN=1" "-DHRT_BQ_WARPS=16 -DHRT_BQ_MIN_CTAS=1 -DHRT_BQ_ALIGN=1"; do
  HRT_EXTRA_NVCC_FLAGS="$v" python -c "import __graft_entry__ as g; g.build(force=True)" || exit 1
  echo "== $v"; for c in closed_form_dof+ ik1_ ik10; do timeout 300 python tools/microbench.py --cases bq --iters 10 --bq-only $c 2>&1 | grep -E "case|rror" | cut -c1-150; done
done
