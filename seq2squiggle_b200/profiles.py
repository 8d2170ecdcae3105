"""Chemistry profiles and option plumbing of ``seq2squiggle predict``.

Mirrors utils.py:129-263 (``get_profile``, ``update_profile``, ``update_config``) and the derived values of
inference.py:348-368 (``dwell_mean = sample_rate / bps``, ``ideal_mode``).  The constants are the reference's
(they are data the drop-in must carry verbatim).
"""
from __future__ import annotations

import logging

logger = logging.getLogger("seq2squiggle")

PROFILE_NAMES = ("dna-r10-prom", "dna-r10-min", "dna-r9-prom", "dna-r9-min", "rna-004-prom", "rna-004-min")

_KEYS = ("digitisation", "sample_rate", "bps", "range", "offset_mean", "offset_std", "median_before_mean",
         "median_before_std")
_TABLE = {
    "dna-r10-min": (8192, 5000, 400, 1536.598389, 13.380569389019, 16.311471649012, 202.15407438804, 13.406139241768),
    "dna-r10-prom": (2048, 5000, 400, 281.345551, -127.5655735, 19.377283387665, 189.87607393756, 15.788097978713),
    "dna-r9-min": (8192, 4000, 450, 1443.030273, 13.7222605, 10.25279688, 200.815801, 20.48933762),
    "dna-r9-prom": (2048, 4000, 450, 748.5801, -237.4102, 14.1575, 214.2890337, 18.0127916),
    "rna-004-min": (8192, 4000, 130, 1437.976685, 12.47686423863, 10.442126577137, 205.08496731088, 8.6671292866233),
    "rna-004-prom": (2048, 4000, 130, 299.432068, -259.421128, 16.010841823643, 189.87607393756, 15.788097978713),
}


def normalise_profile_name(name: str) -> str:
    """The README spells profiles with underscores (``dna_r9_min``); the CLI only knows hyphens."""
    return name.replace("_", "-")


def get_profile(profile: str):
    """utils.py:129-215: a fresh dict per call; unknown names log an error and give None."""
    row = _TABLE.get(profile)
    if row is None:
        logger.error(f"Incorrect value for profile: {profile}")
        return None
    return dict(zip(_KEYS, row))


def update_profile(profile_dict: dict, **kwargs) -> dict:
    """utils.py:218-243: non-None overrides replace profile values; unknown keys only warn."""
    for key, value in kwargs.items():
        if value is not None and key in profile_dict:
            profile_dict[key] = value
        elif key not in profile_dict:
            logger.warning(f"Warning: {key} is not a valid key in the profile")
    return profile_dict


def update_config(profile_name: str, config: dict) -> dict:
    """utils.py:245-263: the k-mer size follows the chemistry."""
    if profile_name.startswith("dna-r10") or profile_name.startswith("rna-004"):
        config["seq_kmer"] = 9
    elif profile_name.startswith("dna-r9"):
        config["seq_kmer"] = 6
    else:
        raise ValueError(f"Unsupported profile name: {profile_name}. Expected 'dna-r10' or 'dna-r9' prefix.")
    return config


def get_seq_kit_and_flow_cell(profile_name: str):
    """signal_io.py:26-60."""
    mapping = {
        "rna-004": ("sqk-rna004", {"prom": "FLO-PRO004RA", "min": "FLO-MIN004RA"}),
        "rna-002": ("sqk-rna002", {"prom": "FLO-PRO002", "min": "FLO-MIN106"}),
        "dna-r10": ("SQK-LSK114", {"prom": "FLO-PRO114", "min": "FLO-MIN114"}),
        "dna-r9": ("SQK-LSK109", {"prom": "FLO-PRO001", "min": "FLO-MIN110"}),
    }
    for prefix, (kit, cells) in mapping.items():
        if profile_name.startswith(prefix):
            key = "prom" if "prom" in profile_name else "min" if "min" in profile_name else None
            if key is None:
                break
            return kit, cells[key]
    raise ValueError(f"Unsupported profile name: {profile_name}")
