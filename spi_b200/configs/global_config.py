"""spi/configs/global_config.py:2-12 (module-level globals mutated at run time, as in the reference)."""
cuda_visible_devices = '0'
device = 'cuda:0'
training_step = 1
log_snapshot = 500
pivotal_training_steps = 0
model_snapshot_interval = 400
run_name = ''

## spi_b200: replay each optimisation iteration as a captured CUDA graph (eager when False or when tests inject noise)
use_cuda_graphs = True

## spi_b200: within an iteration every view shares w_pivot, so the camera-independent tri-plane backbone is evaluated once and
## shared (results identical to re-evaluating it per view, as the reference does); one backward over the summed losses
share_backbone = True
