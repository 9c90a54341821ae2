from gym_ffmp.envs.ffmp import FFMP  # noqa: F401   (/root/reference/src/gym_ffmp/envs/__init__.py:1)
