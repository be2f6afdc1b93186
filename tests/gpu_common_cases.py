"""Model configurations shared by the tests (BASELINE.json configs 1-4, shortened horizons)."""
GOUTSIAS = [0.043, 0.0007, 0.0715, 0.0039, 0.0199264663575241, 0.4791, 0.000199264663575241,
            0.8765e-11, 0.0830269431563506104, 0.5]
# name -> (model file, parameter values, initial state)
#   toggle: test/TestSolverFromFile.f90:28-31; goutsias: examples/transcr6d.f90:23-32,50;
#   repressilator file model: the reference ships no parameter values for it: a=(100,100,100), b=(1,1,1), x0 of
#   examples/repressilator.f90:37
CASES = {
    "toggle": ("toggle.input", [1.0, 100.0, 1.0, 1.0, 100.0, 1.0], [0, 0]),
    "repressilator": ("repressilator.input", [100.0, 100.0, 100.0, 1.0, 1.0, 1.0], [22, 0, 0]),
    "goutsias": ("goutsias.input", GOUTSIAS, [2, 6, 0, 2, 0, 0]),
    "birth_death": ("birth_death.input", [20.0, 1.0], [0]),
    "toggle_test": ("toggle_test.input", [5000.0, 1600.0, 1.0, 1.0], [0, 0]),
}
# golden fixture tag -> (case, t, FSPTOL, KRYTOL, seed)
GOLDEN_RUNS = {
    "toggle_t20": ("toggle", 20.0, 1e-4, 1e-10, 12345),
    "goutsias_t30": ("goutsias", 30.0, 1e-6, 1e-8, 12345),
    "repressilator_t1": ("repressilator", 1.0, 1e-4, 1e-10, 12345),
    "birth_death_t2": ("birth_death", 2.0, 1e-6, 1e-10, 12345),
}
