"""SpiralParameters -- mirrors ft_mpc/controllers/tools/spiral_parameters.py:11-57 (construct-time)."""
import numpy as np


class SpiralParameters:
    def __init__(self, model):
        self.model = model
        self.mass, self.inertia = model.mass, model.inertia
        self.faulty_force = np.asarray(model.faulty_force, float).flatten()
        self.faulty_force_generalized = np.asarray(model.faulty_force_generalized, float).flatten()
        self.D = model.D
        self.beta = np.array([0.0, 0.0, 0.0, 1.0])          # :21  identity => RotFull(beta) = I
        self.calculate_optimal_parameters()

    def calculate_optimal_parameters(self):
        self.omega_des = np.array([0.0, 0.0, 0.6])                                       # :33
        r_dir = np.array([0.0, 1.0, 0.0])
        self.f_virt = 3.5 * r_dir                                                        # :36
        self.compensation_force = np.block([self.f_virt, np.zeros(3)]) - self.faulty_force_generalized   # :37
        self.r = np.linalg.norm(self.f_virt) / (self.mass * np.linalg.norm(self.omega_des) ** 2) * r_dir   # :39
