"""The per-frame loop of sim_full_body_teleop.py:83-129 without the simulator: mocap wire frames in, robot
dof_pos out.

Wire format (mocap_communication/mocap_receiver.py:49-59): a stream of [4-byte big-endian length][pickle of
{'body_pos' (23,3), 'body_quat' (23,4), 'left_hand_pos' (20,3), 'right_hand_pos' (20,3)} float32 arrays].
The decoder is incremental (feed it whatever recv() returned) and unpickles with a whitelist (numpy arrays in
builtin containers only) -- the reference calls pickle.loads on socket data.  Decoding is host-side; the solve is one
request to the resident stream server (or one kernel launch), which reads the wire layout directly: the 23 -> 21 body
rows and the HandNodes finger order are folded into the kernel's index tables."""
import io
import pickle

import numpy as np

from .engine import POS_FULL_BODY_POS, TREE_ROBOT, Engine, engine_from_zero_poses

_ALLOWED = {
    ("numpy.core.multiarray", "_reconstruct"), ("numpy._core.multiarray", "_reconstruct"),
    ("numpy", "ndarray"), ("numpy", "dtype"), ("numpy.core.multiarray", "scalar"), ("numpy._core.multiarray", "scalar"),
    ("collections", "OrderedDict"),
}


class _ArrayUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if (module, name) in _ALLOWED:
            return super().find_class(module, name)
        raise pickle.UnpicklingError(f"mocap frame refers to {module}.{name}: only numpy arrays in builtin containers are accepted")


class WireDecoder:
    """Incremental parser of the length-prefixed pickle stream."""

    def __init__(self, max_frame_bytes=1 << 20):
        self._buf = bytearray()
        self.max_frame_bytes = max_frame_bytes

    def feed(self, chunk: bytes):
        """Append received bytes; returns the list of complete frames (dicts) they finished."""
        self._buf += chunk
        frames = []
        while len(self._buf) >= 4:
            n = int.from_bytes(self._buf[:4], byteorder="big")
            if n > self.max_frame_bytes:
                raise ValueError(f"mocap frame of {n} bytes exceeds the {self.max_frame_bytes}-byte limit (stream out of sync?)")
            if len(self._buf) < 4 + n:
                break
            frames.append(_ArrayUnpickler(io.BytesIO(bytes(self._buf[4:4 + n]))).load())
            del self._buf[:4 + n]
        return frames

    @staticmethod
    def encode(data_dict) -> bytes:
        """The sender's side (for tests and replay): 4-byte big-endian length + pickle."""
        payload = pickle.dumps(data_dict)
        return len(payload).to_bytes(4, byteorder="big") + payload


class TeleopSession:
    """dof = session.step(data_dict) per mocap frame, like the body of the reference's `while True:` loop
    (all-zero body_pos -> the previous dof_pos is repeated, sim_full_body_teleop.py:96,121-122)."""

    def __init__(self, device=0, persistent=True, clamp=False, ik=False, engine=None, mocap_zero_pose=None, target_zero_pose=None,
                 precise_gripper=True):
        """The session owns its context: built from the given zero poses (like the VtrdynFullBodyPosRetargeter the reference's
        loop constructs, sim_full_body_teleop.py:66-72), else from the bundled vtrdyn_full / Hu v5 tables; `engine` hands in a
        caller-owned, already configured one."""
        if engine is not None:
            self._eng = engine
        elif mocap_zero_pose is not None and target_zero_pose is not None:
            self._eng, slot, (parents, _, glob) = engine_from_zero_poses(mocap_zero_pose, target_zero_pose, device)
            if parents.shape[0] != 59:
                raise ValueError(f"the teleop solver needs the 59-joint vtrdyn_full zero pose, got {parents.shape[0]} joints")
            self._eng.configure_pos(POS_FULL_BODY_POS, slot, TREE_ROBOT, glob, precise_gripper)
        else:
            self._eng = Engine(device).set_standard_trees(precise_gripper=precise_gripper)
        self._cfg = (True, bool(persistent), bool(clamp), bool(ik), False, 0)
        self._eng.stream_pos_open(wire_layout=True, persistent=persistent, clamp=clamp, ik=ik)
        self._dof = np.zeros(30, np.float32)
        self.last_dof_pos = np.zeros(30, np.float32)
        self.decoder = WireDecoder()

    def step(self, data_dict):
        body = np.ascontiguousarray(data_dict["body_pos"], dtype=np.float32)
        if body.shape != (23, 3):
            raise ValueError(f"body_pos must be (23, 3), got {body.shape}")
        if np.allclose(body, 0):
            return self.last_dof_pos.copy()
        lh = np.ascontiguousarray(data_dict["left_hand_pos"], dtype=np.float32)
        rh = np.ascontiguousarray(data_dict["right_hand_pos"], dtype=np.float32)
        if lh.shape != (20, 3) or rh.shape != (20, 3):
            raise ValueError("left_hand_pos / right_hand_pos must be (20, 3)")
        if self._eng._pos_stream_cfg != self._cfg:      # another user of this engine re-opened the mailbox in between
            self._eng.stream_pos_open(wire_layout=True, persistent=self._cfg[1], clamp=self._cfg[2], ik=self._cfg[3])
        self._eng.stream_pos_frame(body, lh, rh, None, self._dof)
        self.last_dof_pos = self._dof.copy()
        return self.last_dof_pos.copy()

    def feed_bytes(self, chunk: bytes):
        """Raw socket bytes in -> one dof_pos per mocap frame they completed."""
        return [self.step(d) for d in self.decoder.feed(chunk)]

    def close(self):
        if self._eng._pos_stream_cfg == self._cfg:
            self._eng.stream_pos_close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
