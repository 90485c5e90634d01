"""One-call rollout iteration: actor -> OU noise -> clip*pi/4 -> env step -> replay store -> reset of finished
envs, i.e. the body of the reference training loop (DDPG/trainv2.py:511-531) without ``learn()``, for N
environments, issued as one launch sequence through ``tt_rollout_step``."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import PRECISIONS, check, ptr, stream_ptr


class RolloutEngine:
    def __init__(self, env, agent, store=True, evaluate=False, precision=None, allow_out_of_bar=None):
        self.env, self.agent = env, agent
        self.store, self.evaluate = store, evaluate
        allow = agent.allow_out_of_bar if allow_out_of_bar is None else allow_out_of_bar
        self.precision, _ = _lib.resolve_precision(precision or agent.precision, allow)
        self.L = _lib.load()
        agent.noise.bind_env(env)
        N, dev = env.num_envs, env.device
        self.action = torch.zeros(N, dtype=torch.float32, device=dev)
        self.scaled = torch.zeros(N, dtype=torch.float32, device=dev)
        self.reward = env._reward
        self.done = env._done
        self.iterations = 0
        self._graph, self._graph_k = None, 0

    # ------------------------------------------------------------------ CUDA graph of K iterations
    def capture(self, k=None):
        """Capture K consecutive rollout iterations (2 kernel launches each) into ONE CUDA graph; ``step_graph()`` replays
        it.  Small batches are launch-bound, a graph removes the per-launch host cost.  Every launch of the C ABI goes to the caller's stream and never synchronises, so the sequence is
        capturable as is; the Philox iteration counter lives in device memory.  The ring write position is a by-value
        argument of the captured launches, so K must bring it back to where it started: K * N % mem_size == 0, and K must
        be even (observation double buffer).  Default: the smallest such K."""
        env, m = self.env, self.agent.memory
        N = env.num_envs
        if k is None:
            k = 2
            if self.store:
                import math
                k = m.mem_size // math.gcd(m.mem_size, N)
                k += k & 1
        if k % 2 or (self.store and (k * N) % m.mem_size):
            raise ValueError("capture: K must be even and K * num_envs a multiple of the ring capacity")
        if k > 4096:
            raise ValueError(f"capture: K = {k} iterations is too long a graph; use a ring capacity closer to num_envs")
        self.step()                                   # warm-up outside the capture: one-time attribute / occupancy queries
        self.step()
        torch.cuda.synchronize(env.device)
        g = torch.cuda.CUDAGraph()
        cntr0, cur0, it0 = m.mem_cntr, env._cur, self.iterations
        with torch.cuda.graph(g):
            for _ in range(k):
                self.step()
        # the capture only recorded the launches: roll the host-side bookkeeping back
        m.mem_cntr, env._cur, self.iterations = cntr0, cur0, it0
        self._graph, self._graph_k = g, k
        # what the captured launches carry BY VALUE: the Philox seed, the ring write position, the observation buffer
        # parity, the actor and the optional done-bits buffer.  step_graph() refuses to replay when any of them has moved (reset(seed=...), step() / remember() in between).
        self._graph_key = self._graph_state()
        return k

    def _graph_state(self):
        env, m = self.env, self.agent.memory
        bits = env._done_bits
        return (env.seed_value, (m.mem_cntr % m.mem_size) if self.store else 0, env._cur, id(self.agent.actor),
                None if bits is None else bits.data_ptr())

    def step_graph(self):
        """Replay the captured K iterations.  Returns (obs_next, reward, done) of the LAST of them."""
        if self._graph is None:
            raise RuntimeError("call capture() first")
        if self._graph_state() != self._graph_key:
            raise RuntimeError("step_graph: the captured launches carry the Philox seed, the ring write position, the observation "
                               f"buffer parity and the actor by value; they were {self._graph_key} at capture time and are "
                               f"{self._graph_state()} now (reset(seed=...), or an odd number of step() / a remember() in between) "
                               "-- call capture() again")
        env, m = self.env, self.agent.memory
        self._graph.replay()
        if self.store:
            m.mem_cntr += self._graph_k * env.num_envs
        self.iterations += self._graph_k
        return env._obs_view(env._obs[env._cur]), self.reward, self.done

    def reset(self, seed=None):
        if seed is not None and int(seed) != self.env.seed_value:
            self._graph = None                        # the captured launches carry the old seed
        obs, _ = self.env.reset(seed=seed)
        self.agent.noise.bind_env(self.env)
        self.agent.noise.reset()
        return obs

    def step(self, out=None):
        """One iteration for all envs.  Returns (obs_next, reward, done) views of device buffers; finished envs'
        rows of obs_next already hold the reset observation (their terminal observation went to the ring).
        ``out = (reward, done)``: float32 [N] / uint8 [N] device tensors that receive this iteration's reward and done instead
        of the engine's own buffers -- a caller that reads results back asynchronously alternates two pairs and needs no
        staging copy.  ``out = (reward, done, done_bits)``: additionally the bit-packed flags (``env.set_done_bits``; int32
        [ceil(N / 32)]) -- what such a caller copies to the host instead of the byte array; ``done`` may then be None (the
        engine's own byte buffer)."""
        env, ag = self.env, self.agent
        reward, done = self.reward, self.done
        if out is not None:
            reward = out[0]
            done = out[1] if out[1] is not None else self.done
            for t, dt in ((reward, torch.float32), (done, torch.uint8)):
                if t.dtype != dt or t.numel() != env.num_envs or t.device != self.reward.device or not t.is_contiguous():
                    raise ValueError("step(out=(reward, done)): float32 [N] and uint8 [N] contiguous tensors on the engine's device")
            if len(out) > 2:
                env.set_done_bits(out[2])
        with torch.cuda.device(env.device):
            cur = env._obs[env._cur]
            env._cur ^= 1
            nxt = env._obs[env._cur]
            m = ag.memory
            b = _lib.RolloutBufs(cur.data_ptr(), nxt.data_ptr(), env.ld_obs, ag.noise.x_prev.data_ptr(), self.action.data_ptr(),
                                 self.scaled.data_ptr(), reward.data_ptr(), done.data_ptr(),
                                 m.state_memory.data_ptr() if self.store else None,
                                 m.action_memory.data_ptr() if self.store else None,
                                 m.reward_memory.data_ptr() if self.store else None,
                                 m.new_state_memory.data_ptr() if self.store else None,
                                 m.terminal_memory.data_ptr() if self.store else None, m.mem_size, m.mem_cntr)
            prec = PRECISIONS[self.precision]
            # (two launches: actor + noise + scaling + store of s, a | env step + store of s', r, done + reset + tick)
            check(self.L.tt_rollout_step(env._h, ag.actor._h, C.byref(b), prec, int(self.evaluate), stream_ptr()))
            if self.store:
                m.mem_cntr += env.num_envs
            self.iterations += 1
        return env._obs_view(nxt), reward, done


class AsyncTrainer:
    """BASELINE.json configs[3]: one rollout iteration + one DDPG update (``Agent.learn``) per step, with the update HIDDEN
    under the rollout: the learner's launch sequence (``tt_learn_step``, 14 small dependent kernels), the optional broadcast of
    the new policy (multi-GPU) and its re-pack into the SPARE packed actor run on a side stream while the rollout kernels of
    the same iteration run on the caller's stream.

    * The rollout kernels are persistent and fill every SM, so a dependent chain of small kernels would otherwise only advance
      at kernel boundaries; ``reserve_sms`` SMs are left free for it (``tt_reserve_sms``: -1.4 % rollout throughput per 2 SMs,
      instead of the whole update time serialised behind every iteration).
    * The update of iteration t samples only ring rows completed by iterations < t (``tt_learn_step_window``), and its
      policy is used by iteration t + 1: the usual one-step policy lag of an asynchronous learner.
    * Two packed actors are used alternately; the side stream re-packs into the one the running iteration does not read.

    ``sync`` (``dist.OverlappedSync``): ranks other than ``learner_rank`` receive the flat parameter vector by broadcast."""

    def __init__(self, engine, reserve_sms=2, sync=None, is_learner=True):
        from .agent import CudaActor
        self.eng, self.sync, self.is_learner = engine, sync, bool(is_learner)
        ag = engine.agent
        self.dev = ag.device
        self.ln = ag.learner
        self.flat = self.ln._flat["actor"]
        with torch.cuda.device(self.dev):
            self.actors = [ag.actor, CudaActor(*ag.actor.dims, device=self.dev)]
            self.actors[1].load_state_dict(ag.actor.state_dict())
            self.side = torch.cuda.Stream(device=self.dev)
            self.ev_step, self.ev_pol = torch.cuda.Event(), torch.cuda.Event()
            check(engine.L.tt_reserve_sms(int(reserve_sms)))
        self.reserve_sms = int(reserve_sms)
        self.it = 0
        self.updates = 0

    def close(self):
        with torch.cuda.device(self.dev):
            torch.cuda.current_stream().wait_stream(self.side)
            check(self.eng.L.tt_reserve_sms(0))

    def _window(self):
        """Ring rows that are complete and that the iteration about to run does not write: (begin, count)."""
        m, n = self.eng.agent.memory, self.eng.env.num_envs
        c, cap = m.mem_cntr, m.mem_size
        if c + n <= cap:
            return 0, c
        if n >= cap:
            return 0, 0
        begin = (c + n) % cap
        return begin, (cap - n if c >= cap else max(0, c - begin))

    def step(self, out=None):
        eng, ag = self.eng, self.eng.agent
        with torch.cuda.device(self.dev):
            main = torch.cuda.current_stream()
            if self.it > 0:
                main.wait_event(self.ev_pol)                  # the policy the side stream prepared during the previous iteration
                ag.actor = self.actors[self.it & 1]
            begin, count = self._window()
            if self.it > 0:
                with torch.cuda.stream(self.side):
                    self.side.wait_event(self.ev_step)        # iteration t - 1 has finished: its rows are complete, the spare actor is free
                    if self.is_learner and count >= self.ln.batch:
                        self.ln.learn(repack_into=None, window=(begin, count))
                        self.updates += 1
                    if self.sync is not None and self.sync.on:
                        self.sync.push_policy(self.flat)
                        self.sync.wait_policy()
                    self.actors[(self.it + 1) & 1].load_flat(self.flat)
                    self.ev_pol.record(self.side)
            res = eng.step(out=out)
            self.ev_step.record(main)
            self.it += 1
        return res
