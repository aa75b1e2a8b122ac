"""``python -m azalea_b200.train_cli --config cfg.yml --rundir run`` --
the reference's training command (azalea/train_cli.py:13-53) over GPU
self-play."""
import logging

import click
import torch
import yaml

from . import __version__
from .policy import Policy
from .policy_trainer import train
from .replay_device import DeviceReplayBuffer


@click.command()
@click.option('--config', type=click.Path(exists=True), required=True,
              help='YAML configuration file')
@click.option('--rundir', type=click.Path(), required=True,
              help='Directory to save results from training run')
@click.option('--model', type=click.Path(),
              help='Warm start training from model checkpoint')
@click.option('--replaybuf', type=click.Path(),
              help='Warm start training from replay buffer checkpoint')
@click.option('--max-steps', type=int, default=None,
              help='Stop after this many optimizer steps')
def main(config, rundir, model, replaybuf, max_steps):
    """Train a Hex policy with self-play on the GPU."""
    logging.basicConfig(level=logging.INFO,
                        format='%(asctime)s %(message)s',
                        datefmt='%Y-%m-%d %H:%M:%S')
    logging.info(f'azalea_b200 {__version__}')

    config = yaml.safe_load(open(config))
    if config['device'] == 'auto':
        config['device'] = 'cuda' if torch.cuda.is_available() else 'cpu'

    if model:
        policy = Policy.load(model, device=config['device'])
        logging.info(f'loaded model checkpoint from {model}')
    else:
        policy = Policy()
        policy.initialize(config)

    buf = None
    if replaybuf:
        state = torch.load(replaybuf)
        buf = DeviceReplayBuffer(state['capacity'], state['board_size'], device=config['device'])
        buf.load_state_dict(state)
        logging.info(f'loaded replay buffer checkpoint from {replaybuf}')

    train(policy, config, rundir, replaybuf=buf, max_steps=max_steps)


if __name__ == '__main__':
    main()
