"""LmazeHierCuda -- the reference's two-level planner / actor env over N mazes on one B200.

Host-side mirror of LmazeEnv_v5 (reference gym_lmaze/envs/lmaze_env_v5.py:17-712) and LmazeEnv_v6
(lmaze_env_v6.py, = v5 + safeFovealGoal, :505-523).  Same protocol and method names, batched:

    fov = env.reset()                      # foveal obs  f32 [N, 7, 35, 35]          (lmaze_env_v5.py:102-153)
    loc = env.plannerStep(goals, mask)     # local obs   f32 [N, 4, 35, 35]          (:158-182)
    fov, loc, globalReward, originalReward, globalDone, localDone, fovealGoal, action = env.step(actions)   (:187-292)

`plannerStep` acts on the envs where `mask` is true (None = all): the caller passes the envs whose local
episode just ended (`localDone | globalDone` with autoreset).  `step` advances every env.  All arithmetic
runs in the CUDA kernel behind the C ABI (include/lmaze_b200.h); this class only owns tensors.
"""
import ctypes

import torch

from .. import _abi
from .lmaze_vec_cuda import LmazeVecCuda


class LmazeHierCuda(LmazeVecCuda):
    def __init__(self, num_envs=1, variant="v5", device=None, seed=0, autoreset=True, env_id0=0, random_ball=True,
                 random_goal=True, tune=None, obs_mode="full"):
        if variant not in ("v5", "v6", 5, 6, "lmaze-v5", "lmaze-v6"):
            raise ValueError("LmazeHierCuda is the planner/actor env: variant must be 'v5' or 'v6'")
        self.variant_name = "v6" if variant in ("v6", 6, "lmaze-v6") else "v5"
        super().__init__(num_envs, "v5", device=device, seed=seed, autoreset=autoreset, env_id0=env_id0,
                         random_ball=random_ball, random_goal=random_goal, tune=tune, obs_mode=obs_mode)
        shape = (ctypes.c_int64 * 3)()
        _abi.check(self._lib.lmz_local_obs_shape(self.variant, ctypes.byref(shape)))
        self.full_local_obs_shape = tuple(shape)
        # obs_mode="compact": both observations are the 5x5 crops before the x7 upsample (f32 [N,7,5,5] / [N,4,5,5])
        self.local_obs_shape = tuple(shape) if obs_mode == "full" else (shape[0], 5, 5)
        n, dev = self.num_envs, self.device
        self.loc_obs = torch.zeros((n,) + self.local_obs_shape, dtype=torch.float32, device=dev)
        self.local_reward = torch.zeros(n, dtype=torch.float32, device=dev)
        self._ldone_u8 = torch.zeros(n, dtype=torch.uint8, device=dev)
        self.local_done = self._ldone_u8.view(torch.bool)
        self._loc_err_u8 = torch.zeros(n, dtype=torch.uint8, device=dev)
        self.loc_err = self._loc_err_u8.view(torch.bool)
        self.foveal_goal = torch.full((n,), 12, dtype=torch.uint8, device=dev)      # hot cell of fovealGoal
        ptrs = [_abi.dl(t) for t in (self.loc_obs, self.local_reward, self._ldone_u8, self._loc_err_u8, self.foveal_goal)]
        _abi.check(self._lib.lmz_bind_local_dl(self._h, *[p for p, _ in ptrs]))
        self._local_keepalive = [k for _, k in ptrs]
        self.global_reward, self.global_done, self.fov_obs = self.reward, self.done, self.obs
        self.step_limit, self.foveal_step_limit = 10, 50                           # lmaze_env_v5.py:47-48

    # ------------------------------------------------------------------ protocol
    def plannerStep(self, goals, mask=None):
        """Reference plannerStep (lmaze_env_v5.py:158-182) for the envs where mask is true; returns the local obs.
        mask="auto": the envs that are waiting for their planner (localDone set, or no plannerStep since their
        last reset) -- decided on the device, no host round trip."""
        goals = self._as_actions(goals)
        if isinstance(mask, str):
            if mask != "auto":
                raise ValueError("mask must be a tensor, None or 'auto'")
            pg, kg = _abi.dl(goals)
            _abi.check(self._lib.lmz_planner_step_auto_dl(self._h, pg, self._stream()))
            return self.loc_obs
        if mask is not None:
            mask = torch.as_tensor(mask).to(device=self.device).to(torch.uint8).contiguous()
        pg, kg = _abi.dl(goals)
        pm, km = _abi.dl(mask)
        _abi.check(self._lib.lmz_planner_step_dl(self._h, pg, pm, self._stream()))
        return self.loc_obs

    planner_step = plannerStep

    def expand_local(self, loc=None):
        """Compact local obs f32 [n,4,5,5] -> the reference's [n,4,35,35] image (exact x7 replication)."""
        loc = self.loc_obs if loc is None else loc
        if loc.shape[-1] == self.full_local_obs_shape[-1]:
            return loc
        return loc.repeat_interleave(7, dim=2).repeat_interleave(7, dim=3)

    def goal_plane(self):
        """fovealGoal as the reference returns it: f32 [N, 1, 5, 5] one-hot (lmaze_env_v5.py:165-168)."""
        return torch.nn.functional.one_hot(self.foveal_goal.long(), 25).to(torch.float32).view(-1, 1, 5, 5)

    def step(self, actions, spawn=None, goal_plane=True):
        """Reference step (lmaze_env_v5.py:187-292), batched 8-tuple.  `loc_err` marks the envs where the
        reference's buildLocalObservation would raise IndexError (their local rows are zeros)."""
        actions = self._as_actions(actions)
        spawn = self._as_spawn(spawn)
        pa, ka = _abi.dl(actions)
        ps, ks = _abi.dl(spawn)
        _abi.check(self._lib.lmz_step_dl(self._h, pa, ps, self._stream()))
        return (self.obs, self.loc_obs, self.reward, self.local_reward, self.done, self.local_done,
                self.goal_plane() if goal_plane else self.foveal_goal, actions)

    def safeFovealGoal(self, draws=None):
        """lmaze-v6 safeFovealGoal (lmaze_env_v6.py:505-523): per env a goal 0..24 whose window cell is not a
        wall.  draws: optional int64 [N, K] = the values np.random.randint(0, 25) would return, consumed until
        a non-wall cell comes up; then returns (goals, used).  Default: device RNG, returns goals u8 [N]."""
        goals = torch.empty(self.num_envs, dtype=torch.uint8, device=self.device)
        pg, kg = _abi.dl(goals)
        if draws is None:
            _abi.check(self._lib.lmz_safe_goal_dl(self._h, None, pg, None, self._stream()))
            return goals
        draws = torch.as_tensor(draws).to(device=self.device, dtype=torch.int64).contiguous()
        used = torch.empty(self.num_envs, dtype=torch.int32, device=self.device)
        pd, kd = _abi.dl(draws)
        pu, ku = _abi.dl(used)
        _abi.check(self._lib.lmz_safe_goal_dl(self._h, pd, pg, pu, self._stream()))
        return goals, used

    safe_foveal_goal = safeFovealGoal

    # ------------------------------------------------------------------ state
    def get_state(self):
        """int32 [N, 17]: _abi.STATE_COLS_HIER."""
        out = torch.empty((self.num_envs, _abi.ST_COLS_HIER), dtype=torch.int32, device=self.device)
        p, k = _abi.dl(out)
        _abi.check(self._lib.lmz_get_state_dl(self._h, p, self._stream()))
        return out

    def capture_step(self, actions_buf, spawn_buf=None):
        actions_buf = self._as_actions(actions_buf)
        spawn_buf = self._as_spawn(spawn_buf)
        self.render_obs()
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            self.step(actions_buf, spawn=spawn_buf, goal_plane=False)
        self._graphs = getattr(self, "_graphs", []) + [(graph, actions_buf, spawn_buf)]
        return graph.replay

    def step_host(self, goals_host, actions_host, greward_host, lreward_host, gdone_host, ldone_host):
        """End-to-end planner + actor step with HOST buffers (pinned CPU tensors): H2D goals and actions,
        plannerStep(mask="auto") + step, D2H both rewards and both done flags, stream synchronised on return."""
        n = self.num_envs
        for t in (goals_host, actions_host, greward_host, lreward_host, gdone_host, ldone_host):
            if t.device.type != "cpu" or not t.is_contiguous() or t.numel() != n:
                raise ValueError("step_host takes contiguous CPU tensors of num_envs elements")
        if goals_host.dtype != actions_host.dtype or actions_host.dtype not in (torch.uint8, torch.int32, torch.int64):
            raise ValueError("goals_host / actions_host must share a dtype of uint8/int32/int64")
        if greward_host.dtype != torch.float32 or lreward_host.dtype != torch.float32:
            raise ValueError("reward buffers must be float32")
        ad = (torch.uint8, torch.int32, torch.int64).index(actions_host.dtype)
        _abi.check(self._lib.lmz_hier_step_host(
            self._h, goals_host.data_ptr(), actions_host.data_ptr(), ad, greward_host.data_ptr(),
            lreward_host.data_ptr(), gdone_host.data_ptr(), ldone_host.data_ptr(), self._stream()))
        return greward_host, lreward_host, gdone_host, ldone_host

    def rollout(self, T, goals=None, actions=None):
        """T fused planner + actor steps, no per-step observations (lmz_hier_rollout): per step plannerStep(goals[t])
        for the envs waiting for their planner, then step(actions[t]).  goals / actions: [T, N] of one integer dtype,
        or both None for device-side random ones.  Returns (globalReward f32, originalReward f32, globalDone bool,
        localDone bool), each [T, N]."""
        T = int(T)
        if (goals is None) != (actions is None):
            raise ValueError("give both goals and actions, or neither")
        if goals is not None:
            goals, actions = self._as_actions(goals), self._as_actions(actions)
            if goals.dtype != actions.dtype:
                goals = goals.to(actions.dtype)
        n, dev = self.num_envs, self.device
        gr = torch.empty((T, n), dtype=torch.float32, device=dev)
        lr = torch.empty((T, n), dtype=torch.float32, device=dev)
        gd = torch.empty((T, n), dtype=torch.uint8, device=dev)
        ld = torch.empty((T, n), dtype=torch.uint8, device=dev)
        ptrs = [_abi.dl(t) for t in (goals, actions, gr, lr, gd, ld)]
        _abi.check(self._lib.lmz_hier_rollout_dl(self._h, T, *[p for p, _ in ptrs], self._stream()))
        return gr, lr, gd.view(torch.bool), ld.view(torch.bool)

    def _unsupported(self, *a, **k):
        raise NotImplementedError("not available for the planner/actor env (lmaze-v5/v6)")

    set_window = render_window = initState = host_pipeline = _unsupported
