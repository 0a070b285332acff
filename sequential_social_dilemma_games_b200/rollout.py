"""Random-action rollouts rendered to frames / video: the tool side of SURVEY.md 8f-4.

Mirrors the reference's `rollout.py` (`Controller.rollout` / `render_rollout`, rollout.py:30-110) and the video helpers of
`utility_funcs.py:8-56`, on top of the batched engine: the episode is stepped by the fused CUDA kernel, the full-map frames
(`MapEnv.map_to_colors()` of `get_map_with_agents()`, map_env.py:280-339) come from `ssd_render_map` for every env of the
batch at once, and only the frames of the env being filmed cross PCIe.  Differences from upstream, both deliberate:
`agents[0].action_space.n` raises in this fork of the reference (rollout.py:64, agent.py:56), here the number of actions is
the game's (8 / 9); and the 'pretty' mode writes PNG frames with OpenCV instead of matplotlib (not installed in this image).
"""
import os
import shutil

import numpy as np
import torch

from .batched import BatchedSSDEnv
from .config import make_config
from .video import make_video_from_image_dir, make_video_from_rgb_imgs, save_img


class Controller(object):
    """rollout.py:30-110 over `num_envs` environments resident on the GPU; env `film` is the one whose frames are kept."""

    def __init__(self, env_name='cleanup', num_agents=5, num_envs=1, device="cuda:0", seed=0, film=0):
        if env_name not in ('harvest', 'cleanup'):
            raise ValueError('Error! Not a valid environment type')
        self.env_name = env_name
        self.cfg = make_config(env_name, num_agents=num_agents)
        self.env = BatchedSSDEnv(self.cfg, num_envs, device=device, seed=seed)
        self.film = int(film)
        self._gen = torch.Generator(device=self.env.device).manual_seed(seed)
        self.env.reset()

    def rollout(self, horizon=50, save_path=None):
        """`horizon` steps with uniform random actions.  Returns (rewards of agent-0, observations of agent-0 as the float64
        arrays the reference returns, full-map uint8 frames [H, W, 3]) of the filmed env; PNG frames go to `save_path`."""
        B, N = self.env.num_envs, self.cfg.num_agents
        rewards, observations, full_obs = [], [], []
        for i in range(horizon):
            actions = torch.randint(0, self.cfg.num_actions, (B, N), generator=self._gen, device=self.env.device, dtype=torch.int8)
            obs, rew = self.env.step(actions)
            frame = self.env.render_map()[self.film].cpu().numpy()
            if save_path is not None:
                save_img(frame, save_path, 'frame' + str(i).zfill(6) + '.png')
            full_obs.append(frame)
            observations.append((obs[self.film, 0].cpu().numpy().astype(np.float64) - 128.0) / 255.0)   # map_env.py:199
            rewards.append(int(rew[self.film, 0]))
        return rewards, observations, full_obs

    def render_rollout(self, horizon=50, path=None, render_type='pretty', fps=8):
        """A rollout as <path>/<env>_trajectory.mp4; 'pretty' goes through PNG frames, 'fast' straight from the arrays."""
        if path is None:
            path = os.path.join(os.getcwd(), 'videos')
        os.makedirs(path, exist_ok=True)
        video_name = self.env_name + '_trajectory'
        if render_type == 'pretty':
            image_path = os.path.join(path, 'frames')
            os.makedirs(image_path, exist_ok=True)
            self.rollout(horizon=horizon, save_path=image_path)
            out = make_video_from_image_dir(path, image_path, fps=fps, video_name=video_name)
            shutil.rmtree(image_path)
            return out
        _, _, full_obs = self.rollout(horizon=horizon)
        return make_video_from_rgb_imgs([f[:, :, ::-1] for f in full_obs], path, fps=fps, video_name=video_name)  # OpenCV writes BGR

    def close(self):
        self.env.close()


def main(argv=None):
    import argparse
    ap = argparse.ArgumentParser(description="random-action rollout of one environment to a video (reference: rollout.py)")
    ap.add_argument('--vid_path', default=os.path.join(os.getcwd(), 'videos'))
    ap.add_argument('--env', default='cleanup', choices=['cleanup', 'harvest'])
    ap.add_argument('--render_type', default='pretty', choices=['pretty', 'fast'])
    ap.add_argument('--fps', type=int, default=8)
    ap.add_argument('--horizon', type=int, default=50)
    args = ap.parse_args(argv)
    c = Controller(env_name=args.env)
    print(c.render_rollout(horizon=args.horizon, path=args.vid_path, render_type=args.render_type, fps=args.fps))


if __name__ == '__main__':
    main()
