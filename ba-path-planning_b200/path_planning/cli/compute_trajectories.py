"""``compute-trajectories``: one randomized scenario, reference defaults
(src/path_planning/cli/compute_trajectories.py:9-99: N=10, T=100 s, h=0.2 s -> K=500, R=0.8 m,
200 x 200 m), solved on the GPU."""

import time

from ..scenarios.position_generator import generate_positions
from ..solvers.scp import SCP


def main():
    print("------ WOW Fleet Collision-Free 2D Trajectory Generation ------")
    n_vehicles, time_horizon, time_step, min_distance = 10, 100, 0.2, 0.8
    space_dims = [0, 0, 200, 200]
    print("Configuration:")
    print(f"  Number of vehicles: {n_vehicles}")
    print(f"  Time horizon: {time_horizon} s")
    print(f"  Time step: {time_step} s")
    print(f"  Minimum margin: {min_distance} m")
    print(f"  Space dimensions: {space_dims} m")
    print()
    solver = SCP(n_vehicles=n_vehicles, time_horizon=time_horizon, time_step=time_step,
                 min_distance=min_distance, space_dims=space_dims)
    initial_positions, final_positions = generate_positions(n_vehicles, min_distance)
    print(f"Successfully generated positions for {n_vehicles} vehicles")
    solver.set_initial_states(initial_positions)
    solver.set_final_states(final_positions)
    print("Generating trajectories...")
    start_time = time.time()
    try:
        solver.generate_trajectories(max_iterations=15)
        end_time = time.time()
        print("\nTrajectory generation complete!")
        print(f"Total computation time: {end_time - start_time:.3f} seconds")
        print(f"Number of time steps: {solver.K}")
        print(f"Total trajectory duration: {solver.T} seconds")
        print("\nVisualizing 2D trajectories...")
        solver.visualize_trajectories(show_animation=True)
        print("\nVisualizing time snapshots")
        solver.visualize_time_snapshots(num_snapshots=5)
    except Exception as e:  # the reference prints and swallows (compute_trajectories.py:98-99)
        print(f"Error during trajectory generation: {e}")


if __name__ == "__main__":
    main()
