"""Helpers shared by the golden-trace tests."""
import glob
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    meta = json.loads(str(z["meta"]))
    return meta, {k: z[k] for k in z.files if k != "meta"}


def env_kwargs(meta):
    return dict(num_agents=meta["num_agents"], max_num_timesteps=meta["max_num_timesteps"],
                communication_on=meta["communication_on"], num_communication=meta["num_communication"],
                ego_led=meta["ego_led"], fow_radius=meta["fow_radius"],
                ego_config=meta["ego_config"], partner_config=meta["partner_config"])
