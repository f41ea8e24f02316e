"""Electrode-solver parity cases shared by the golden generator, the oracle tests and the GPU tests."""
import cases


def electrode_cases():
    three = cases.blobs3(40, seed=40)
    return {
        "el_blobs48":     ("ElectrodeSolver", cases.blobs(48, 0.5, seed=48), {}, {}),
        "el_blobs48_per": ("PeriodicElectrodeSolver", cases.blobs(48, 0.5, seed=48), {}, {}),
        "el_rand":        ("ElectrodeSolver", cases.random_img((24, 20, 16), 0.7, 11), {}, {}),
        "el_rand_per":    ("PeriodicElectrodeSolver", cases.random_img((24, 20, 16), 0.7, 11), {}, {}),
        "el_odd":         ("ElectrodeSolver", cases.odd_random(8), {}, {"iter_limit": 300}),
        "el_odd_per":     ("PeriodicElectrodeSolver", cases.odd_random(8), {}, {"iter_limit": 300}),
        "el_labels":      ("ElectrodeSolver", three, {"conductive_label": 2, "reactive_label": 0, "spacing": 0.5}, {}),
        "el_labels_per":  ("PeriodicElectrodeSolver", three, {"conductive_label": 1, "reactive_label": 2, "omega": 1.8}, {"conv_crit": 1e-3}),
        "el_batch":       ("ElectrodeSolver", cases.stacked_blobs(2, 32, 100), {}, {}),
        "el_blobs64":     ("ElectrodeSolver", cases.blobs(64, 0.5, seed=64), {}, {"conv_crit": 1e-3}),
    }
