"""Minimal configargparse: argparse plus `--config file` whose lines are `key = value` (`# comments`, `[a, b, c]` lists
for action="append" options, empty / 0 / false for store_true flags).  Command-line arguments override the file."""
import argparse
import sys


class ArgumentParser(argparse.ArgumentParser):
    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self._config_dest = None

    def add_argument(self, *names, **kw):
        if kw.pop("is_config_file", False):
            act = super().add_argument(*names, **kw)
            self._config_dest = act.dest
            return act
        return super().add_argument(*names, **kw)

    def _file_args(self, path):
        out = []
        actions = {a.dest: a for a in self._actions}
        for line in open(path):
            line = line.split("#", 1)[0].strip()
            if not line or "=" not in line:
                continue
            key, val = (t.strip() for t in line.split("=", 1))
            act = actions.get(key)
            if act is None:
                continue
            flag = "--" + key
            if isinstance(act, argparse._StoreTrueAction):
                if val.lower() not in ("", "0", "false", "none"):
                    out.append(flag)
            elif val.startswith("["):
                for item in val.strip("[]").split(","):
                    if item.strip():
                        out += [flag, item.strip()]
            else:
                out += [flag, val]
        return out

    def parse_args(self, args=None, namespace=None):
        args = list(sys.argv[1:] if args is None else args)
        pre = argparse.ArgumentParser(add_help=False)
        pre.add_argument("--config", default=None)
        known, _ = pre.parse_known_args(args)
        file_args = self._file_args(known.config) if known.config else []
        return super().parse_args(file_args + args, namespace)
