"""Minimal stand-in for the reference's hydra/omegaconf config surface (conf/config.yaml + groups):
`load_config(["mlp.hidden_size=256", "masking=Pruning", "quant=none", "+img.index=3"])`.
hydra is not a dependency: groups are plain YAML files under conf/<group>/<option>.yaml, overrides use
hydra's `group=option` / `dotted.key=value` syntax, `${a.b}` interpolations are resolved at load time.
`mlp.width` is accepted as an alias of `mlp.hidden_size` (README-era name)."""
import os
import re

import yaml

CONF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "conf")
GROUPS = ("img", "mlp", "optim", "masking", "quant", "entropy_coding")


class Config(dict):
    """dict with attribute access and omegaconf-like .get(); nested dicts are wrapped on access."""

    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError as e:
            raise AttributeError(k) from e
        return v

    def __setattr__(self, k, v):
        self[k] = v

    @staticmethod
    def wrap(obj):
        if isinstance(obj, dict):
            return Config({k: Config.wrap(v) for k, v in obj.items()})
        if isinstance(obj, list):
            return [Config.wrap(v) for v in obj]
        return obj


def _parse_value(text):
    return yaml.safe_load(text)


def _set_dotted(cfg, dotted, value):
    keys = dotted.split(".")
    node = cfg
    for k in keys[:-1]:
        if k not in node or not isinstance(node[k], dict):
            node[k] = Config()
        node = node[k]
    node[keys[-1]] = value


def _get_dotted(cfg, dotted):
    node = cfg
    for k in dotted.split("."):
        node = node[k]
    return node


def _resolve(cfg, node=None):
    node = cfg if node is None else node
    for k, v in list(node.items()):
        if isinstance(v, dict):
            _resolve(cfg, v)
        elif isinstance(v, str):
            for _ in range(8):
                m = re.search(r"\$\{([^}:]+)\}", v) if isinstance(v, str) else None
                if not m:
                    break
                ref = _get_dotted(cfg, m.group(1))
                v = ref if m.group(0) == v else v.replace(m.group(0), str(ref))
            node[k] = v


def load_config(overrides=(), conf_dir=CONF_DIR):
    with open(os.path.join(conf_dir, "config.yaml")) as f:
        root = yaml.safe_load(f)
    choices = dict(root.pop("defaults"))
    plain = []
    for ov in overrides:
        ov = ov.lstrip("+")
        key, _, val = ov.partition("=")
        if key in GROUPS:
            choices[key] = val
        else:
            plain.append((key, val))
    cfg = Config.wrap(root)
    for group, option in choices.items():
        path = os.path.join(conf_dir, group, f"{option}.yaml")
        if not os.path.exists(path):
            raise FileNotFoundError(f"no config option {group}={option} ({path})")
        with open(path) as f:
            cfg[group] = Config.wrap(yaml.safe_load(f) or {})
    for key, val in plain:
        if key == "mlp.width":
            key = "mlp.hidden_size"
        _set_dotted(cfg, key, Config.wrap(_parse_value(val)))
    _resolve(cfg)
    return cfg
