# A/B timing of the edge step on one GPU: forward kernels (HGNN_FWD_PP=0 two-CTA, 1 ping-pong);
# forward in inference mode (no stash), forward in training mode, forward + backward
for ws in ${WS_LIST:-0 1}; do
  for mode in ${MODES:-infer fwd full}; do
    echo "WS=$ws mode=$mode"; HGNN_FWD_PP=$ws python profiles/edge_step_once.py 1000000 12 $mode
  done
done
