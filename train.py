"""train.py - same flags as the reference CLI (train.py:134-181); builds the env,
the algorithm and the Trainer on the B200 kernels.  Launch one process per GPU
with torchrun to shard the rollout environments across GPUs."""
import argparse
import datetime
import os

import numpy as np
import yaml


def train(args):
    print(f"> Running train.py {args}")
    import torch
    import torch.distributed as dist
    from dgppo_b200.algo import make_algo
    from dgppo_b200.env import make_env
    from dgppo_b200.trainer.trainer import Trainer

    if "RANK" in os.environ and int(os.environ.get("WORLD_SIZE", "1")) > 1:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
    np.random.seed(args.seed)
    env = make_env(env_id=args.env, num_agents=args.num_agents, num_obs=args.obs, n_rays=args.n_rays,
                   full_observation=args.full_observation)
    env_test = make_env(env_id=args.env, num_agents=args.num_agents, num_obs=args.obs, n_rays=args.n_rays,
                        full_observation=args.full_observation)
    algo = make_algo(
        algo=args.algo, env=env, node_dim=env.node_dim, edge_dim=env.edge_dim, state_dim=env.state_dim,
        action_dim=env.action_dim, n_agents=env.num_agents, cost_weight=args.cost_weight,
        cbf_weight=args.cbf_weight, actor_gnn_layers=args.actor_gnn_layers, Vl_gnn_layers=args.Vl_gnn_layers,
        Vh_gnn_layers=args.Vh_gnn_layers, rnn_layers=args.rnn_layers, lr_actor=args.lr_actor, lr_Vl=args.lr_Vl,
        lr_Vh=args.lr_Vh, max_grad_norm=2.0, alpha=args.alpha, cbf_eps=args.cbf_eps, seed=args.seed,
        batch_size=args.batch_size, use_rnn=not args.no_rnn, use_lstm=args.use_lstm, coef_ent=args.coef_ent,
        rnn_step=args.rnn_step, gamma=0.99, clip_eps=args.clip_eps, lagr_init=args.lagr_init,
        lr_lagr=args.lr_lagr, train_steps=args.steps, cbf_schedule=not args.no_cbf_schedule,
        cost_schedule=args.cost_schedule)
    start_time = datetime.datetime.now().strftime("%m%d%H%M%S")
    log_dir = f"{args.log_dir}/{args.env}/{args.algo}/seed{args.seed}_{start_time}"
    if args.name is not None:
        log_dir = f"{log_dir}_{args.name}"
    os.makedirs(log_dir, exist_ok=True)
    train_params = {"run_name": args.name or start_time, "training_steps": args.steps,
                    "eval_interval": args.eval_interval, "eval_epi": args.eval_epi,
                    "save_interval": args.save_interval}
    trainer = Trainer(env=env, env_test=env_test, algo=algo, log_dir=log_dir, n_env_train=args.n_env_train,
                      n_env_test=args.n_env_test, seed=args.seed, params=train_params, save_log=not args.debug)
    if not args.debug:
        with open(f"{log_dir}/config.yaml", "w") as f:
            yaml.dump(args, f)
            yaml.dump(algo.config, f)
    trainer.train()


# The reference CLI (train.py:134-181), flag for flag: (flags, type or None for a switch, default | REQUIRED)
REQUIRED = object()
CLI = [
    # what to run
    (("--env",), str, REQUIRED), (("-n", "--num-agents"), int, REQUIRED), (("--algo",), str, REQUIRED),
    (("--obs",), int, REQUIRED),
    # run control, algorithm coefficients
    (("--seed",), int, 0), (("--steps",), int, 200000), (("--name",), str, None), (("--debug",), None, False),
    (("--cost-weight",), float, 0.0), (("--n-rays",), int, 32), (("--full-observation",), None, False),
    (("--clip-eps",), float, 0.25), (("--lagr-init",), float, 0.5), (("--lr-lagr",), float, 1e-7),
    (("--cbf-weight",), float, 1.0), (("--cbf-eps",), float, 1e-2), (("--alpha",), float, 10.0),
    (("--no-cbf-schedule",), None, False), (("--cost-schedule",), None, False), (("--no-rnn",), None, False),
    # networks
    (("--actor-gnn-layers",), int, 2), (("--Vl-gnn-layers",), int, 2), (("--Vh-gnn-layers",), int, 1),
    (("--lr-actor",), float, 3e-4), (("--lr-Vl",), float, 1e-3), (("--lr-Vh",), float, 1e-3),
    (("--rnn-layers",), int, 1), (("--use-lstm",), None, False), (("--coef-ent",), float, 1e-2),
    (("--rnn-step",), int, 16),
    # sizes, logging
    (("--n-env-train",), int, 128), (("--batch-size",), int, 16384), (("--n-env-test",), int, 32),
    (("--log-dir",), str, "./logs"), (("--eval-interval",), int, 50), (("--eval-epi",), int, 1),
    (("--save-interval",), int, 50),
]


def build_parser() -> argparse.ArgumentParser:
    parser = argparse.ArgumentParser(description="DGPPO training on the B200 rollout kernels")
    for flags, typ, default in CLI:
        if typ is None:
            parser.add_argument(*flags, action="store_true", default=default)
        elif default is REQUIRED:
            parser.add_argument(*flags, type=typ, required=True)
        else:
            parser.add_argument(*flags, type=typ, default=default)
    return parser


def main():
    train(build_parser().parse_args())


if __name__ == "__main__":
    main()
