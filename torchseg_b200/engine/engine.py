"""Engine / State with the reference API (/root/reference/furnace/engine/engine.py:23-163): argparse injection
(-d / -c / --local_rank), WORLD_SIZE detection, process-group init, state registry, checkpoint
save / link / restore in the reference's on-disk format.

Differences (documented in INTEGRATION.md): the backend is NCCL when CUDA is present and gloo otherwise
(CPU tests), LOCAL_RANK from the environment is honoured (torchrun), `-d ''` no longer raises, and a
world of one process under torchrun still counts as "distributed" when TSB_FORCE_DISTRIBUTED=1.
"""
import argparse
import os
import os.path as osp
import time
from collections import OrderedDict

import torch
import torch.distributed as dist

from .logger import get_logger
from .version import __version__
from ..utils.pyt_utils import load_model, parse_devices, extant_file, link_file, ensure_dir

logger = get_logger()


class State(object):
    """engine.py:23-35"""

    def __init__(self):
        self.epoch = 0
        self.iteration = 0
        self.dataloader = None
        self.model = None
        self.optimizer = None

    def register(self, **kwargs):
        for k, v in kwargs.items():
            assert k in ['epoch', 'iteration', 'dataloader', 'model', 'optimizer']
            setattr(self, k, v)


class Engine(object):
    def __init__(self, custom_parser=None, argv=None):
        self.version = __version__
        logger.info("PyTorch Version {}, Furnace(B200) Version {}".format(torch.__version__, self.version))
        self.state = State()
        self.devices = None
        self.distributed = False
        self.local_rank = 0
        self.world_size = 1
        if custom_parser is None:
            self.parser = argparse.ArgumentParser()
        else:
            assert isinstance(custom_parser, argparse.ArgumentParser)
            self.parser = custom_parser
        self.inject_default_parser()
        self.args = self.parser.parse_args(argv)
        self.continue_state_object = self.args.continue_fpath

        if 'WORLD_SIZE' in os.environ:
            self.distributed = int(os.environ['WORLD_SIZE']) > 1 or os.environ.get('TSB_FORCE_DISTRIBUTED') == '1'
        if self.distributed:
            self.local_rank = int(os.environ.get('LOCAL_RANK', self.args.local_rank))
            self.world_size = int(os.environ['WORLD_SIZE'])
            use_cuda = torch.cuda.is_available()
            if use_cuda:
                torch.cuda.set_device(self.local_rank)
            if not dist.is_initialized():
                dist.init_process_group(backend="nccl" if use_cuda else "gloo", init_method='env://')
            self.devices = [i for i in range(self.world_size)]
        else:
            self.devices = parse_devices(self.args.devices)

    def inject_default_parser(self):
        p = self.parser
        p.add_argument('-d', '--devices', default='', help='set data parallel training')
        p.add_argument('-c', '--continue', type=extant_file, metavar="FILE", dest="continue_fpath",
                       help='continue from one certain checkpoint')
        p.add_argument('--local_rank', default=0, type=int, help='process rank on node')

    def register_state(self, **kwargs):
        self.state.register(**kwargs)

    def update_iteration(self, epoch, iteration):
        self.state.epoch = epoch
        self.state.iteration = iteration

    def save_checkpoint(self, path):
        """engine.py:89-115: {'model' (no 'module.' prefix), 'optimizer', 'epoch', 'iteration'}"""
        logger.info("Saving checkpoint to file {}".format(path))
        t0 = time.time()
        model_sd = OrderedDict()
        for k, v in self.state.model.state_dict().items():
            key = k[7:] if k.split('.')[0] == 'module' else k
            # on-disk weights are plain contiguous NCHW fp32, whatever the in-memory (KRSC) layout
            model_sd[key] = v.detach().contiguous().cpu() if torch.is_tensor(v) else v
        state_dict = {'model': model_sd, 'optimizer': self.state.optimizer.state_dict(), 'epoch': self.state.epoch,
                      'iteration': self.state.iteration}
        t1 = time.time()
        torch.save(state_dict, path)
        logger.info("Save checkpoint to file {}, Time usage:\n\tprepare snapshot: {}, IO: {}".format(
            path, t1 - t0, time.time() - t1))

    def save_and_link_checkpoint(self, snapshot_dir, log_dir, log_dir_link):
        """engine.py:117-126"""
        ensure_dir(snapshot_dir)
        if not osp.exists(log_dir_link):
            link_file(log_dir, log_dir_link)
        current = osp.join(snapshot_dir, 'epoch-{}.pth'.format(self.state.epoch))
        self.save_checkpoint(current)
        link_file(current, osp.join(snapshot_dir, 'epoch-last.pth'))

    def restore_checkpoint(self):
        """engine.py:128-152"""
        t0 = time.time()
        tmp = torch.load(self.continue_state_object, map_location=torch.device('cpu'), weights_only=False)
        t1 = time.time()
        has_module = any(k.startswith('module.') for k in self.state.model.state_dict().keys())
        self.state.model = load_model(self.state.model, tmp['model'], has_module)
        self.state.optimizer.load_state_dict(tmp['optimizer'])
        self.state.epoch = tmp['epoch'] + 1
        self.state.iteration = tmp['iteration']
        del tmp
        logger.info("Load checkpoint from file {}, Time usage:\n\tIO: {}, restore snapshot: {}".format(
            self.continue_state_object, t1 - t0, time.time() - t1))

    def __enter__(self):
        return self

    def __exit__(self, type, value, tb):
        if torch.cuda.is_available():
            torch.cuda.empty_cache()
        if type is not None:
            logger.warning("A exception occurred during Engine initialization, give up running process")
            return False
