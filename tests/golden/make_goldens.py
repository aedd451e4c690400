"""Writes the golden fixtures from the literals in the reference's own tests (file:line in README.md)."""
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
ddim = {
    "config": {"num_train_timesteps": 1000, "beta_start": 0.0001, "beta_end": 0.02, "beta_schedule": "linear", "clip_sample": True},
    "steps_offset_1_set_timesteps_5": [801, 601, 401, 201, 1],
    "variance": [[0, 0, 0.0], [420, 400, 0.14771], [980, 960, 0.32460], [487, 486, 0.00979], [999, 998, 0.02]],
    "variance_atol": 1e-5,
    "full_loop": [
        {"config": {}, "sum": 172.0067, "mean": 0.223967},
        {"config": {"prediction_type": "v_prediction"}, "sum": 52.5302, "mean": 0.0684},
        {"config": {"set_alpha_to_one": True, "beta_start": 0.01}, "sum": 149.8295, "mean": 0.1951},
        {"config": {"set_alpha_to_one": False, "beta_start": 0.01}, "sum": 149.0784, "mean": 0.1941},
    ],
    "full_loop_with_noise": {"t_start": 8, "sum": 354.5418, "mean": 0.4616},
    "sum_atol": 1e-2, "mean_atol": 1e-3, "num_inference_steps": 10,
}
sinus = {
    "embedding_dim": 64, "timesteps": 128, "slice": [[23, 26], [47, 50]], "atol": 0.01,
    "cases": [
        {"kwargs": {"downscale_freq_shift": 1, "flip_sin_to_cos": False},
         "values": [0.9646, 0.9804, 0.9892, 0.9615, 0.9787, 0.9882, 0.9582, 0.9769, 0.9872]},
        {"kwargs": {"downscale_freq_shift": 0, "flip_sin_to_cos": True},
         "values": [0.3019, 0.228, 0.1716, 0.3146, 0.2377, 0.179, 0.3272, 0.2474, 0.1864]},
        {"kwargs": {"scale": 1000},
         "values": [-0.9801, -0.9464, -0.9349, -0.3952, 0.8887, -0.9709, 0.5299, -0.2853, -0.9927]},
    ],
}
euler = {  # ppdiffusers/tests/schedulers/test_scheduler_euler.py:25-33 (config), :63-200 (full loops, 10 steps)
    "config": {"num_train_timesteps": 1100, "beta_start": 0.0001, "beta_end": 0.02, "beta_schedule": "linear"},
    "num_inference_steps": 10,
    "full_loop": [
        {"config": {}, "sum": 10.0807, "mean": 0.0131},
        {"config": {"prediction_type": "v_prediction"}, "sum": 0.0002, "mean": 2.2676e-06},
        {"config": {"use_karras_sigmas": True}, "sum": 124.52299499511719, "mean": 0.16213932633399963},
    ],
    "full_loop_with_noise": {"t_start": 8, "sum": 57062.9023, "mean": 74.3007},
    "sum_atol": 1e-2, "mean_atol": 1e-3,
}
dpm = {  # ppdiffusers/tests/schedulers/test_scheduler_dpm_multi.py:33-52 (config), :116-131 (full_loop), :229-284
    "config": {"num_train_timesteps": 1000, "beta_start": 0.0001, "beta_end": 0.02, "beta_schedule": "linear", "solver_order": 2,
               "prediction_type": "epsilon", "algorithm_type": "dpmsolver++", "solver_type": "midpoint",
               "lower_order_final": False, "euler_at_final": False},
    "num_inference_steps": 10,
    "full_loop": [
        {"config": {}, "mean": 0.3301},
        {"config": {"prediction_type": "v_prediction"}, "mean": 0.2251},
        {"config": {"prediction_type": "v_prediction", "use_karras_sigmas": True}, "mean": 0.2096},
        {"config": {"prediction_type": "v_prediction", "use_lu_lambdas": True}, "mean": 0.1554},
    ],
    "full_loop_with_noise": {"t_start": 5, "sum": 318.4111, "mean": 0.4146},
    "sum_atol": 1e-2, "mean_atol": 1e-3,
}
lcm = {  # ppdiffusers/tests/schedulers/test_scheduler_lcm.py:28-38 (config), :239-247 (one-step full loop, RNG-free)
    "config": {"num_train_timesteps": 1000, "beta_start": 0.00085, "beta_end": 0.012, "beta_schedule": "scaled_linear",
               "prediction_type": "epsilon"},
    "one_step": {"sum": 18.7097, "mean": 0.0244}, "atol": 1e-3,
    "one_step_timesteps": [999],  # scheduling_lcm.py:430-437 with original_inference_steps = 50
    "ten_step_timesteps": [999, 899, 799, 699, 599, 499, 399, 299, 199, 99],
}
json.dump(lcm, open(os.path.join(HERE, "lcm_goldens.json"), "w"), indent=1)
json.dump(dpm, open(os.path.join(HERE, "dpm_multistep_goldens.json"), "w"), indent=1)
json.dump(euler, open(os.path.join(HERE, "euler_goldens.json"), "w"), indent=1)
json.dump(ddim, open(os.path.join(HERE, "ddim_goldens.json"), "w"), indent=1)
json.dump(sinus, open(os.path.join(HERE, "sinusoid_goldens.json"), "w"), indent=1)
