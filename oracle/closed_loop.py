"""ORACLE (test infrastructure) -- the reference's receding-horizon loop, CasadiOptimizer.optimize()
(/root/reference/MPC_Planner/optimizer.py:562-643), noise-free, with oracle/ipm.py in place of IPOPT."""
import numpy as np

from . import ipm, nlp


def optimize(sc, N, verbose=False):
    """sc: scenario namespace (mpc_b200.load_scenario).  Returns (traj_s[T,5], u[T,2]) like optimizer.py:637-643 (Q12)."""
    T = sc.iter_length
    init_state = np.array(sc.x0, float)
    current = init_state.copy()
    u0 = np.zeros((N, 2))
    next_traj = np.tile(current, (N + 1, 1))          # optimizer.py:581 (Q4)
    next_states = next_traj.copy()
    traj, u_c = [], []
    for i in range(T):
        d = nlp.make_nlp(N, sc.dt, sc.weights_setting, next_traj, sc.static_obstacle)
        r = ipm.solve(d, nlp.pack(u0, next_states))
        if r["status"] != 1:
            raise RuntimeError(f"oracle failed at MPC step {i}: status {r['status']} kkt {r['kkt']}")
        U, X = nlp.split(r["w"], N)
        u_c.append(U[0].copy())
        current = nlp.euler_step(current, U[0], sc.dt)                     # shift_movement, optimizer.py:649-650
        u0 = np.concatenate([U[1:], U[-1:]])                               # optimizer.py:652
        next_states = np.concatenate([X[1:], X[-1:]])                      # optimizer.py:653
        next_states[0] = current
        next_traj = nlp.reference_window(i, current, N, T, sc.reference_path, sc.orientation, sc.desired_velocity)
        traj.append(current.copy())
        if verbose:
            print(i, r["iters"], U[0])
    traj_s = np.array(traj)
    traj_s = np.insert(traj_s, 0, init_state, axis=0)[:-1]
    return traj_s, np.array(u_c)
