"""Trains a PPO agent on an example env (the offline counterpart of the reference's
examples/train_agent.py, which drives rl_zoo3).

    python examples/train_agent.py -e DiscreteSteps-v0 -a ppo [--num-envs 64] [--rollouts 4]
    torchrun --nproc-per-node 8 examples/train_agent.py -e DiscreteSteps-v0 -a ppo --num-envs 4096
"""

import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    parser = argparse.ArgumentParser(prog="python train_agent.py")
    parser.add_argument("-e", "--env", required=True, choices=["DiscreteSteps-v0"])
    parser.add_argument("-a", "--algo", required=True, choices=["ppo"])
    parser.add_argument("--num-envs", type=int, default=None, help="default: n_envs of the yml")
    parser.add_argument("--rollouts", type=int, default=2)
    parser.add_argument("--max-minibatches", type=int, default=None)
    parser.add_argument("--device-env", action="store_true",
                        help="step DeviceVectorDiscreteSteps: env step, buffer and normalisation on the GPU")
    args = parser.parse_args()

    import yaml

    from examples import ppo

    conf = os.path.join(os.path.dirname(os.path.abspath(__file__)), f"{args.algo}_tuned.yml")
    cfg = ppo.PPOConfig.from_yml(conf, args.env)
    with open(conf) as f:
        n_envs = args.num_envs or yaml.safe_load(f)[args.env]["n_envs"]
    history = ppo.train(num_envs=n_envs, rollouts=args.rollouts, config=cfg,
                        max_minibatches=args.max_minibatches, log=lambda e: print(json.dumps(e)),
                        device_env=args.device_env)
    return history


if __name__ == "__main__":
    main()
