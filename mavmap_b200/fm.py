"""Host-side stand-in for the reference FeatureManager *data layout*.

The BA reads and writes the reference's `FeatureManager` public maps in place
(src/fm/feature_management.h:189-230; Ceres optimises that storage directly,
src/base3d/bundle_adjustment.cc:243-247,269-270).  This class exposes the same
public members with the same names and the same 1-based id allocation
(src/fm/feature_management.cc:42,53,64,79) so that the Python parity tests read like
code written against the reference.  Track merge / dedup logic
(feature_management.cc:107-245) is out of scope; `add_track` is a
direct way to record a triangulated point with its observations.
"""
import numpy as np


class FeatureManager:
    def __init__(self):
        self.points3D = {}             # {P3D_ID: xyz(3)}
        self.points3D_tri = {}         # {P3D_ID: bool}
        self.points2D = {}             # {P2D_ID: xy(2)}
        self.point2D_to_point3D = {}   # {P2D_ID: P3D_ID}
        self.point2D_to_image = {}     # {P2D_ID: IMG_ID}
        self.image_to_points2D = {}    # {IMG_ID: [P2D_ID, ...]}
        self.point3D_to_points2D = {}  # {P3D_ID: [P2D_ID, ...]}
        self.rvecs = {}                # {IMG_ID: rvec(3)}
        self.tvecs = {}                # {IMG_ID: tvec(3)}
        self.image_to_camera = {}      # {IMG_ID: CAM_ID}
        self.camera_params = {}        # {CAM_ID: [fx, fy, cx, cy, ..., model_code]}
        self._num_cameras = self._num_images = self._num_points2D = self._num_points3D = 0

    def get_num_cameras(self): return self._num_cameras
    def get_num_images(self): return self._num_images
    def get_num_points2D(self): return self._num_points2D
    def get_num_points3D(self): return self._num_points3D

    def get_point2D_idx(self, point2D_id):
        return point2D_id - self.image_to_points2D[self.point2D_to_image[point2D_id]][0]

    def add_point2D(self, image_id, xy):
        self._num_points2D += 1
        pid = self._num_points2D
        self.points2D[pid] = np.asarray(xy, dtype=np.float64).copy()
        self.image_to_points2D[image_id].append(pid)
        self.point2D_to_image[pid] = image_id
        return pid

    def add_point3D(self):
        self._num_points3D += 1
        pid = self._num_points3D
        self.points3D[pid] = np.zeros(3)
        self.points3D_tri[pid] = False
        self.point3D_to_points2D[pid] = []
        return pid

    def add_camera(self, params):
        """params = model parameters followed by the model code (sequential_mapper.cc:960-965)."""
        self._num_cameras += 1
        self.camera_params[self._num_cameras] = [float(p) for p in params]
        return self._num_cameras

    def add_image(self, camera_id, points2D=None):
        self._num_images += 1
        iid = self._num_images
        self.image_to_camera[iid] = camera_id
        self.rvecs[iid] = np.zeros(3)
        self.tvecs[iid] = np.zeros(3)
        self.image_to_points2D[iid] = []
        if points2D is not None:
            for xy in points2D:
                self.add_point2D(iid, xy)
        return iid

    def set_point3D(self, point3D_id, xyz):
        self.points3D[point3D_id] = np.asarray(xyz, dtype=np.float64).copy()
        self.points3D_tri[point3D_id] = True

    def set_pose(self, image_id, rvec, tvec):
        self.rvecs[image_id] = np.asarray(rvec, dtype=np.float64).copy()
        self.tvecs[image_id] = np.asarray(tvec, dtype=np.float64).copy()

    def add_track(self, xyz, observations):
        """Record a 3-D point observed at [(image_id, point2D_idx), ...]."""
        pid = self.add_point3D()
        self.set_point3D(pid, xyz)
        for image_id, idx in observations:
            p2 = self.image_to_points2D[image_id][idx]
            self.point2D_to_point3D[p2] = pid
            self.point3D_to_points2D[pid].append(p2)
        return pid
