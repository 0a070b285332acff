"""Frames -> files: the video helpers of the reference's utility_funcs.py:8-56 (PNG frames, mp4 through OpenCV), used by
rollout.py.  No CUDA here: importable (and tested) on a machine without a GPU."""
import os

import numpy as np


def make_video_from_rgb_imgs(rgb_arrs, vid_path, video_name='trajectory', fps=5, format="mp4v", resize=(640, 480)):
    """A list of uint8 RGB frames [H, W, 3] -> <vid_path>/<video_name>.mp4 (utility_funcs.py:27-56: nearest-neighbour resize,
    frames handed to the writer as they are).  Returns the path of the file."""
    import cv2
    os.makedirs(vid_path, exist_ok=True)
    video_path = os.path.join(vid_path, video_name + '.mp4')
    if resize is not None:
        width, height = resize
    else:
        height, width = rgb_arrs[0].shape[:2]
    video = cv2.VideoWriter(video_path, cv2.VideoWriter_fourcc(*format), float(fps), (width, height))
    for image in rgb_arrs:
        image = np.ascontiguousarray(image, dtype=np.uint8)
        if resize is not None:
            image = cv2.resize(image, resize, interpolation=cv2.INTER_NEAREST)
        video.write(image)
    video.release()
    return video_path


def make_video_from_image_dir(vid_path, img_folder, video_name='trajectory', fps=5):
    """Every .png of a directory, in name order, into one video (utility_funcs.py:13-24)."""
    import cv2
    images = sorted(f for f in os.listdir(img_folder) if f.endswith(".png"))
    return make_video_from_rgb_imgs([cv2.imread(os.path.join(img_folder, f)) for f in images], vid_path,
                                    video_name=video_name, fps=fps)


def save_img(rgb_arr, path, name, scale=16):
    """One frame as a PNG, every cell `scale` x `scale` pixels (utility_funcs.py:8-10 uses matplotlib for this)."""
    import cv2
    img = np.ascontiguousarray(np.asarray(rgb_arr, dtype=np.uint8)[:, :, ::-1])   # OpenCV writes BGR
    img = cv2.resize(img, (img.shape[1] * scale, img.shape[0] * scale), interpolation=cv2.INTER_NEAREST)
    cv2.imwrite(os.path.join(path, name), img)
