"""Same attribute names as the reference's config.py:1-20, plus the knobs the B200 path adds."""


class Config(object):
    epoch = 5000
    batch_size = 8
    learning_rate = 0.002

    sigma = 1.0

    cuda = True
    gpu_cnt = 4            # reference: nn.DataParallel device count; here informational (one process per GPU)

    async_loading = True
    pin_memory = True

    root_path = '.'
    data_path = 'data/dataset'
    checkpoint_dir = 'model'
    checkpoint_file = 'checkpoint.pth.tar'
    summary_dir = 'board'

    pretraining_step_size = 220

    # --- additions ---
    bucket_mb = 64         # NCCL gradient bucket size
    vae_head = False
    packed_array_path = None  # directory of flat memory-mappable bit arrays (tools/pack_dataset.py --arrays); wins if it exists
    packed_data_path = None   # directory of bit-packed .npz items (tools/pack_dataset.py); used when it exists
    packed_input = False   # loader emits bit-packed batches (data/packed.py): 32x less host->device traffic
    refiner = False        # graph/refiner.py with the one-line shape fix applied after the decoder (graph/model.py:31,41)
    micro_bars = 0         # > 0: steps over more bars run as gradient-accumulated chunks of this size (trainer.py)
    gan_schedule = 'with_gan'   # agent.barGen_with_gan: 'with_gan' (agent/barGen_with_gan.py) or 'horovod' (agent/barGen_horovod.py)
    synthetic = False      # explicit opt-in: train on random SyntheticBars when no dataset directory exists (smoke runs)
