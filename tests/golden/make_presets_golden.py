"""Dumps the reference's hparam_presets dict to presets_golden.json (build container only)."""
import importlib.util
import json
from pathlib import Path

spec = importlib.util.spec_from_file_location('ref_presets', '/root/reference/hparam_presets.py')
mod = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mod)
Path(__file__).with_name('presets_golden.json').write_text(json.dumps(mod.hparam_presets, indent=1, sort_keys=True))
