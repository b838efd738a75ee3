"""Scenario generation (input producer of the hot path; host side, microseconds).

``generate_positions`` keeps the reference's contract
(src/path_planning/scenarios/position_generator.py:44-75): N starts on four
corner circles, N goals on a central diamond (90 %) or the circles (10 %),
pairwise spacing >= min_distance, stdlib ``random`` consumed in the reference's
order so that ``random.seed(s)`` reproduces the reference's scenarios bit for
bit (pinned in tests/).  Plot/diagnostic helpers of the reference module are
out of scope.

``generate_positions_large`` is NEW: the reference generator cannot place
N >= 100 vehicles and simply scaling its layout makes goals
unreachable under |v| <= 2 m/s, so large synthetic scenarios use bounded-travel
random start/goal pairs instead (documented in DESIGN.md).
"""

from __future__ import annotations

import math
import random

import numpy as np

ARENA = 20.0
_CIRCLE_R = 2.5
_CORNERS = np.array([[3.5, 3.5], [16.5, 3.5], [3.5, 16.5], [16.5, 16.5]])
_MID = np.array([ARENA / 2, ARENA / 2])
_REACH = 6.0 / np.sqrt(2)  # centre -> diamond vertex for a 6 m side
_DIAMOND = np.array(
    [
        [_MID[0], _MID[1] + _REACH],
        [_MID[0] + _REACH, _MID[1]],
        [_MID[0], _MID[1] - _REACH],
        [_MID[0] - _REACH, _MID[1]],
    ]
)


class _Sampler:
    """Draws candidates; each method consumes ``random`` exactly like the reference helper it mirrors."""

    @staticmethod
    def corner_circle():
        c = _CORNERS[random.randint(0, 3)]
        phi = random.uniform(0, 2 * np.pi)
        return c + _CIRCLE_R * np.array([np.cos(phi), np.sin(phi)])

    @staticmethod
    def diamond_edge():
        e = random.randint(0, 3)
        a, b = _DIAMOND[e], _DIAMOND[(e + 1) % 4]
        s = random.uniform(0, 1)
        return a + s * (b - a)

    @staticmethod
    def goal():
        return _Sampler.diamond_edge() if random.random() < 0.9 else _Sampler.corner_circle()


def _place(n, draw, min_distance, max_attempts, what):
    placed = []
    for _ in range(max_attempts):
        if len(placed) >= n:
            break
        cand = draw()
        if all(np.linalg.norm(cand - q) >= min_distance for q in placed):
            placed.append(cand)
    if len(placed) < n:
        raise ValueError(f"Could not generate enough {what} positions.")
    return np.array(placed)


def generate_positions(n_vehicles, min_distance=0.4, max_attempts=1000):
    """(initial (N,2), final (N,2)); same signature and semantics as the reference."""
    starts = _place(n_vehicles, _Sampler.corner_circle, min_distance, max_attempts, "initial")
    goals = _place(n_vehicles, _Sampler.goal, min_distance, max_attempts, "final")
    return starts, goals


def generate_positions_large(n_vehicles, min_distance=0.8, time_horizon=20.0, vel_limit=2.0,
                             area_per_vehicle=16.0, spacing_factor=1.25, max_attempts=None):
    """Synthetic scenario for N beyond the reference generator's reach.

    Square arena of side sqrt(N * area_per_vehicle); starts uniform with pairwise
    spacing >= spacing_factor * min_distance; each goal = start + a displacement of
    length U(0.5, 1) * d_max in a uniform direction, d_max = 0.4 * vel_limit *
    time_horizon (a rest-to-rest minimum-acceleration move of length d peaks at
    1.5 d / T, so d_max keeps the velocity box inactive), re-drawn until it lies
    inside the arena and keeps the goal spacing.  Uses stdlib ``random`` like the
    reference generator.  Returns (initial, final, space_dims).
    """
    side = math.sqrt(n_vehicles * area_per_vehicle)
    gap = spacing_factor * min_distance
    d_max = 0.4 * vel_limit * time_horizon
    max_attempts = max_attempts or 200 * n_vehicles
    starts = np.empty((n_vehicles, 2))
    goals = np.empty((n_vehicles, 2))
    n = 0
    for _ in range(max_attempts):
        if n >= n_vehicles:
            break
        s = np.array([random.uniform(1.0, side - 1.0), random.uniform(1.0, side - 1.0)])
        if n and np.min(np.hypot(*(starts[:n] - s).T)) < gap:
            continue
        ok = False
        for _ in range(20):
            ang = random.uniform(0.0, 2 * math.pi)
            d = random.uniform(0.5, 1.0) * d_max
            g = s + d * np.array([math.cos(ang), math.sin(ang)])
            if not (1.0 <= g[0] <= side - 1.0 and 1.0 <= g[1] <= side - 1.0):
                continue
            if n and np.min(np.hypot(*(goals[:n] - g).T)) < gap:
                continue
            ok = True
            break
        if not ok:
            continue
        starts[n], goals[n] = s, g
        n += 1
    if n < n_vehicles:
        raise ValueError("Could not generate enough positions.")
    return starts, goals, [0.0, 0.0, side, side]


def print_distance_analysis(initial_positions, final_positions):
    """Distance report of a start/goal set (reference position_generator.py:173-205): smallest pairwise spacing over
    both sets and the longest straight-line trip; same printed block and returned keys.  The per-trajectory counterpart
    (sampled and continuous-time separation of a SOLVED scenario) is `path_planning.analysis.check_trajectories`."""
    a = np.asarray(initial_positions, dtype=float)
    b = np.asarray(final_positions, dtype=float)

    def closest(points):
        if len(points) < 2:
            return float("inf")
        i, j = np.triu_indices(len(points), 1)
        return float(np.linalg.norm(points[i] - points[j], axis=1).min())

    trips = np.linalg.norm(b - a, axis=1)
    report = dict(global_min_distance=min(closest(a), closest(b)), longest_path=trips.max(), longest_vehicle=trips.argmax())
    bar = "=" * 40
    print(f"\n{bar}\nDISTANCE SUMMARY\n{bar}")
    print(f"Global minimum distance: {report['global_min_distance']:.3f} m")
    print(f"Longest path traveled:  {report['longest_path']:.3f} m (Vehicle {report['longest_vehicle']})")
    print(bar + "\n")
    return report
