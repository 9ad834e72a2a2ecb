#!/usr/bin/env bash
# End-to-end check of the `unicore-train` CLI on 2 GPUs: fp16 + deferred overflow check + EMA + NVLink DDP,
# checkpoint save / resume, then bf16 + stochastic rounding + gradient accumulation.  Run from the repo root.
set -e
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
rm -rf gpurun_out/cli_ck && mkdir -p gpurun_out/cli_ck
COMMON="--user-dir examples/bert --task synthetic_mlm --loss masked_lm --arch bert_base --encoder-layers 4 --encoder-embed-dim 256 --encoder-ffn-embed-dim 1024 --encoder-attention-heads 4 --synthetic-vocab-size 2048 --synthetic-seq-len 128 --synthetic-num-samples 4096 --max-seq-len 128 --optimizer adam --adam-betas (0.9,0.98) --clip-norm 1.0 --lr-scheduler polynomial_decay --lr 5e-4 --warmup-updates 5 --total-num-update 200 --batch-size 16 --log-format simple --log-interval 5 --num-workers 0 --weight-decay 0.01 --seed 3 --save-dir gpurun_out/cli_ck --tmp-save-dir gpurun_out/cli_ck --save-interval-updates 20 --validate-interval-updates 20 --ema-decay 0.999 --ddp-backend b200 --pin-memory"
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 unicore_cli/train.py $COMMON "${@:2}"; }
echo "== fp16 + deferred overflow, 20 updates"
run 29601 --fp16 --fp16-init-scale 128 --deferred-overflow-check --max-update 20 2>&1 | grep -E "train_inner|valid \||Saved|rror|Traceback" | tail -8
echo "== resume to 30 updates"
run 29602 --fp16 --fp16-init-scale 128 --deferred-overflow-check --max-update 30 2>&1 | grep -E "Loaded|train_inner|rror|Traceback" | tail -6
ls gpurun_out/cli_ck | head
rm -rf gpurun_out/cli_ck && mkdir -p gpurun_out/cli_ck
echo "== bf16 + stochastic rounding, update-freq 2"
run 29603 --bf16 --bf16-sr --update-freq 2 --max-update 10 2>&1 | grep -E "train_inner|rror|Traceback" | tail -5
rm -rf gpurun_out/cli_ck
