"""Joint-name tables and index ranges shared by the fitters.

Values are facts about the SMPL skeleton / the reference's conventions
(/root/reference/keypoints2body/core/constants.py:1-71); they have to agree
with the reference for the API to be a drop-in.
"""

# SMPL 24-joint body layout (+ vertex-picked face/feet extras used by SMPL24).
JOINT_MAP = dict(
    MidHip=0, LHip=1, RHip=2, spine1=3, LKnee=4, RKnee=5, spine2=6, LAnkle=7,
    RAnkle=8, spine3=9, LFoot=10, RFoot=11, Neck=12, LCollar=13, Rcollar=14,
    Head=15, LShoulder=16, RShoulder=17, LElbow=18, RElbow=19, LWrist=20,
    RWrist=21, LHand=22, RHand=23,
    Nose=24, REye=25, LEye=26, REar=27, LEar=28, LHeel=31, RHeel=34,
)

SMPL_IDX = range(24)

# AMASS uses the first 22 SMPL joints (no hand joints).
AMASS_JOINT_MAP = {k: v for k, v in JOINT_MAP.items() if v < 22}

AMASS_IDX = range(22)
AMASS_SMPL_IDX = range(22)

# Opaque model-joint index blocks for dict ("GENERIC") observations
# (reference constants.py:65-71).
SMPLX_BODY_IDX = range(22)
SMPLX_LEFT_HAND_IDX = range(25, 46)
SMPLX_RIGHT_HAND_IDX = range(46, 67)
SMPLX_FACE_IDX_START = 67

# Body-pose entries (0-based inside the 69-D body_pose) carrying the elbow/knee
# bending prior and the sign of the exponent (reference losses.py:13-21).
ANGLE_PRIOR_IDX = (52, 55, 9, 12)
ANGLE_PRIOR_SIGN = (1.0, -1.0, -1.0, -1.0)

# Joints whose confidence `fix_foot` raises (reference api/sequence.py:124-128).
FIX_FOOT_IDX = (7, 8, 10, 11)
FIX_FOOT_CONF = 1.5
