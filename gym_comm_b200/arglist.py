"""Env-config surface of the reference (``arglist.py:4-121``, ``env_args.json``, ``spread/*.json``).

Same keys, same meaning.  Where the reference forgets a key and crashes (SURVEY Appendix C:
``ego_config`` / ``partner_config`` missing -> ``KeyError`` in arglist.py:104; ``CAN_MOVE`` missing
-> ``KeyError`` in gym_comm/envs/overcooked_env.py:254) the drop-in supplies the obvious default
``{CAN_MOVE: true, ALLERGIC: false, BLIND: false}``.
"""
from __future__ import annotations

import argparse
import json

DEFAULT_AGENT_CONFIG = {"CAN_MOVE": True, "ALLERGIC": False, "BLIND": False}


def _agent_config(d):
    out = dict(DEFAULT_AGENT_CONFIG)
    out.update(d or {})
    return out


def namespace_from_dict(args: dict) -> argparse.Namespace:
    """dict with the JSON keys -> Namespace with the attribute names the env reads
    (argparse dest names of arglist.py:38-94)."""
    if "level" not in args or "num_agents" not in args:
        raise KeyError("env config needs 'level' and 'num_agents' (arglist.py:40-41: required)")
    return argparse.Namespace(
        level=args["level"],
        num_agents=int(args["num_agents"]),
        max_num_timesteps=int(args.get("max_num_timesteps", 100)),       # arglist.py:42 default
        max_num_subtasks=int(args.get("max_num_subtasks", 14)),
        seed=int(args.get("seed", 1)),
        hyperparams=dict(args.get("hyperparams", {})),
        total_timesteps=int(args.get("total_timesteps", 20000000)),
        record_interval=int(args.get("record_interval", 500)),
        log=bool(args.get("log", False)),
        notes=args.get("notes", "XXX notes"),
        wandb=bool(args.get("wandb", False)),
        communication_on=bool(args.get("communication_on", False)),
        num_communication=int(args.get("num_communication", 10)),         # arglist.py:71-74
        ego_led=bool(args.get("ego_led", False)),
        fow_radius=int(args.get("fow_radius", 2)),
        ego_config=_agent_config(args.get("ego_config")),
        partner_config=_agent_config(args.get("partner_config")),
        play=False, record=False, with_image_obs=False,
        model1=None, model2=None, model3=None, model4=None,
    )


def create_arglist(json_path: str) -> argparse.Namespace:
    """Reference: ``arglist.create_arglist(json_path)``."""
    with open(json_path, "r") as f:
        return namespace_from_dict(json.load(f))


def normalize(arglist) -> argparse.Namespace:
    """Accept a Namespace built by the reference's own parser (or a dict) and fill the gaps."""
    d = dict(arglist) if isinstance(arglist, dict) else dict(vars(arglist))
    return namespace_from_dict(d)
