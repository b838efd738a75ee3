"""ORACLE / TEST INFRASTRUCTURE ONLY -- restatement of the reference's scenario generator.

Follows /root/reference/src/path_planning/scenarios/position_generator.py:
layout constants :17-40, generate_positions :44-75, sampling helpers :235-248.
Randomness is the stdlib ``random`` module (position_generator.py:10), consumed
in exactly the reference's order, so ``random.seed(s)`` gives bit-identical
scenarios (pinned in tests/test_oracle_vs_reference.py and tests/golden/).
"""

from __future__ import annotations

import random

import numpy as np

BOX = 20.0
RADIUS = 5.0 / 2.0
CENTERS = np.array([[3.5, 3.5], [16.5, 3.5], [3.5, 16.5], [16.5, 16.5]])
_C = np.array([BOX / 2, BOX / 2])
_HALF_DIAG = 6.0 / np.sqrt(2)
DIAMOND = np.array(
    [[_C[0], _C[1] + _HALF_DIAG], [_C[0] + _HALF_DIAG, _C[1]], [_C[0], _C[1] - _HALF_DIAG], [_C[0] - _HALF_DIAG, _C[1]]]
)


def _on_circle(center):  # position_generator.py:236-238
    ang = random.uniform(0, 2 * np.pi)
    return center + RADIUS * np.array([np.cos(ang), np.sin(ang)])


def _on_diamond():  # position_generator.py:241-245
    e = random.randint(0, 3)
    a, b = DIAMOND[e], DIAMOND[(e + 1) % 4]
    t = random.uniform(0, 1)
    return a + t * (b - a)


def _far_enough(p, placed, dmin):  # position_generator.py:248
    return all(np.linalg.norm(p - o) >= dmin for o in placed)


def generate_positions(n_vehicles, min_distance=0.4, max_attempts=1000):
    """position_generator.py:44-75."""
    starts, tries = [], 0
    while len(starts) < n_vehicles and tries < max_attempts:
        cand = _on_circle(CENTERS[random.randint(0, 3)])
        if _far_enough(cand, starts, min_distance):
            starts.append(cand)
        tries += 1
    if len(starts) < n_vehicles:
        raise ValueError("Could not generate enough initial positions.")
    goals, tries = [], 0
    while len(goals) < n_vehicles and tries < max_attempts:
        if random.random() < 0.9:
            cand = _on_diamond()
        else:
            cand = _on_circle(CENTERS[random.randint(0, 3)])
        if _far_enough(cand, goals, min_distance):
            goals.append(cand)
        tries += 1
    if len(goals) < n_vehicles:
        raise ValueError("Could not generate enough final positions.")
    return np.array(starts), np.array(goals)
