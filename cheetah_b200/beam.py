"""Beam state containers (mirror of cheetah/particles/{particle_beam,parameter_beam}.py).

Mirrored: the constructor contract (particle_beam.py:60-106, parameter_beam.py:27-60), the
coordinate views, the survival-weighted moments and everything derived from them (means, sigmas,
the 15 named covariances, emittances, Twiss and dispersion functions: beam.py:262-557,
particle_beam.py:1699-1951, parameter_beam.py:643-749), the conversions between the two beam types,
the SI phase-space round trip and the Gaussian set-up generators.  For a ParticleBeam on a CUDA
device the first and second moments come from ONE pass of the fused covariance kernel
(``second_moments``); file converters and plotting stay the reference's job (SURVEY.md 2, rows 8
and 10).
"""

from __future__ import annotations

import torch
from torch import nn

from .species import SPEED_OF_LIGHT, Species

COORDINATES = ("x", "px", "y", "py", "tau", "p")
# named second moments (beam.py:343-430) -> index pairs into the 6 x 6 covariance
COVARIANCES = {
    "cov_xpx": (0, 1), "cov_ypy": (2, 3), "cov_taup": (4, 5), "cov_xp": (0, 5),
    "cov_pxp": (1, 5), "cov_yp": (2, 5), "cov_pyp": (3, 5), "cov_xy": (0, 2),
    "cov_xpy": (0, 3), "cov_xtau": (0, 4), "cov_pxy": (1, 2), "cov_pxpy": (1, 3),
    "cov_pxtau": (1, 4), "cov_ytau": (2, 4), "cov_pytau": (3, 4),
}


class Beam(nn.Module):
    """Quantities derived from the first and second moments (cheetah/particles/beam.py:262-557);
    subclasses provide ``mu_*``, ``sigma_*`` and ``_covariance(i, j)``."""

    def _covariance(self, i: int, j: int) -> torch.Tensor:
        raise NotImplementedError

    def _plane(self, position: int):
        # sigma^2 and covariances of one transverse plane with the dispersive part removed
        # (beam.py:441-486 for x, :497-536 for y)
        sigma_p2 = getattr(self, "sigma_p").square()
        u, pu = COORDINATES[position], COORDINATES[position + 1]
        cov_up = self._covariance(position, 5)
        cov_pup = self._covariance(position + 1, 5)
        uu = getattr(self, f"sigma_{u}").square() - cov_up.square() / sigma_p2
        pp = getattr(self, f"sigma_{pu}").square() - cov_pup.square() / sigma_p2
        upu = self._covariance(position, position + 1) - cov_up * cov_pup / sigma_p2
        return uu, pp, upu

    def _emittance(self, position: int) -> torch.Tensor:
        uu, pp, upu = self._plane(position)
        return (uu * pp - upu.square()).clamp_min(torch.finfo(uu.dtype).tiny).sqrt()

    def _projected_emittance(self, position: int) -> torch.Tensor:
        u, pu = COORDINATES[position], COORDINATES[position + 1]
        return (
            getattr(self, f"sigma_{u}").square() * getattr(self, f"sigma_{pu}").square()
            - self._covariance(position, position + 1).square()
        ).sqrt()

    emittance_x = property(lambda self: self._emittance(0))
    emittance_y = property(lambda self: self._emittance(2))
    projected_emittance_x = property(lambda self: self._projected_emittance(0))
    projected_emittance_y = property(lambda self: self._projected_emittance(2))
    normalized_emittance_x = property(
        lambda self: self.emittance_x * self.relativistic_beta * self.relativistic_gamma)
    normalized_emittance_y = property(
        lambda self: self.emittance_y * self.relativistic_beta * self.relativistic_gamma)
    beta_x = property(lambda self: self._plane(0)[0] / self.emittance_x)
    beta_y = property(lambda self: self._plane(2)[0] / self.emittance_y)
    alpha_x = property(lambda self: -self._plane(0)[2] / self.emittance_x)
    alpha_y = property(lambda self: -self._plane(2)[2] / self.emittance_y)
    dispersion_x = property(lambda self: self._covariance(0, 5) / self.sigma_p.square())
    dispersion_px = property(lambda self: self._covariance(1, 5) / self.sigma_p.square())
    dispersion_y = property(lambda self: self._covariance(2, 5) / self.sigma_p.square())
    dispersion_py = property(lambda self: self._covariance(3, 5) / self.sigma_p.square())

    @property
    def relativistic_gamma(self) -> torch.Tensor:
        return self.energy / self.species.mass_eV

    @property
    def relativistic_beta(self) -> torch.Tensor:
        # cheetah/particles/beam.py:328-336
        gamma = self.relativistic_gamma
        beta = torch.ones_like(gamma)
        nonzero = gamma.abs() > 0
        beta[nonzero] = (1 - gamma[gamma > 0].square().reciprocal()).sqrt()
        return beta

    @property
    def p0c(self) -> torch.Tensor:
        return self.relativistic_beta * self.relativistic_gamma * self.species.mass_eV


for _name, (_i, _j) in COVARIANCES.items():
    setattr(Beam, _name, property(lambda self, i=_i, j=_j: self._covariance(i, j)))


class ParticleBeam(Beam):
    """Beam of macroparticles: ``particles (..., N, 7)`` rows ``[x, px, y, py, tau, delta, 1]``."""

    def __init__(
        self,
        particles: torch.Tensor,
        energy: torch.Tensor,
        particle_charges: torch.Tensor | None = None,
        survival_probabilities: torch.Tensor | None = None,
        s: torch.Tensor | None = None,
        species: Species | None = None,
        device: torch.device | None = None,
        dtype: torch.dtype | None = None,
    ) -> None:
        super().__init__()
        assert (
            particles.shape[-2] > 0 and particles.shape[-1] == 7
        ), "Particle vectors must be 7-dimensional."
        device = device if device is not None else particles.device
        dtype = dtype if dtype is not None else particles.dtype
        factory_kwargs = {"device": device, "dtype": dtype}
        self.species = species if species is not None else Species("electron", **factory_kwargs)
        self.register_buffer("particles", particles)
        self.register_buffer("energy", energy)
        self.register_buffer(
            "particle_charges",
            particle_charges
            if particle_charges is not None
            else torch.full(
                (particles.shape[-2],), float(self.species.charge_coulomb), **factory_kwargs
            ),
        )
        self.register_buffer(
            "survival_probabilities",
            survival_probabilities
            if survival_probabilities is not None
            else torch.ones(particles.shape[-2], **factory_kwargs),
        )
        self.register_buffer("s", s if s is not None else torch.tensor(0.0, **factory_kwargs))
        # None = unknown; the tracker checks once whether particles[..., 6] == 1 and caches it
        self._unit_seventh: bool | None = None

    @classmethod
    def _from_tracking(cls, particles, energy, particle_charges, survival_probabilities, s,
                       species, unit_seventh=None) -> "ParticleBeam":
        """Constructor used by the tracker for outgoing beams: same object as ``__init__`` builds,
        without its argument checks and ``register_buffer`` bookkeeping (~60 us per call)."""
        beam = cls.__new__(cls)
        nn.Module.__init__(beam)
        beam._modules["species"] = species
        buffers = beam._buffers
        buffers["particles"] = particles
        buffers["energy"] = energy
        buffers["particle_charges"] = particle_charges
        buffers["survival_probabilities"] = survival_probabilities
        buffers["s"] = s
        object.__setattr__(beam, "_unit_seventh", unit_seventh)
        return beam

    # ---- set-up helpers (host-side convenience, not part of the accelerated path) --------
    @classmethod
    def from_distribution(
        cls,
        mu: torch.Tensor,
        cov: torch.Tensor,
        num_particles: int = 100_000,
        energy: torch.Tensor | None = None,
        total_charge: torch.Tensor | None = None,
        s: torch.Tensor | None = None,
        species: Species | None = None,
        device: torch.device | None = None,
        dtype: torch.dtype | None = None,
        generator: torch.Generator | None = None,
    ) -> "ParticleBeam":
        """Gaussian beam with mean ``mu (..., 6)`` and covariance ``cov (..., 6, 6)``.

        As in the reference (particle_beam.py:357-431, ``match_distribution_moments``,
        utils/statistics.py:91-150) ONE standard-normal sample ``(N, 6)`` is whitened to exactly
        zero mean and unit covariance and then mapped by the Cholesky factor of every target
        covariance, so the sample moments of every vector entry equal ``(mu, cov)`` exactly.  The
        sampling runs in float64 on the host (``generator`` for reproducibility)."""
        dtype = dtype if dtype is not None else torch.get_default_dtype()
        factory_kwargs = {"device": device, "dtype": dtype}
        species = species if species is not None else Species("electron", **factory_kwargs)
        energy = energy if energy is not None else torch.tensor(1e8, **factory_kwargs)
        if total_charge is None:
            total_charge = species.charge_coulomb * num_particles
        total_charge = torch.as_tensor(total_charge, **factory_kwargs)
        particle_charges = (
            torch.ones((*total_charge.shape, num_particles), **factory_kwargs)
            * total_charge.unsqueeze(-1)
            / num_particles
        )
        mu64 = torch.as_tensor(mu).detach().to("cpu", torch.float64)
        cov64 = torch.as_tensor(cov).detach().to("cpu", torch.float64)
        standard = torch.randn(num_particles, 6, dtype=torch.float64, generator=generator)
        if num_particles > 6:  # whiten the sample (needs a non-singular sample covariance)
            centred = standard - standard.mean(dim=0, keepdim=True)
            sample_factor = torch.linalg.cholesky(centred.mT @ centred / (num_particles - 1))
            standard = torch.linalg.solve_triangular(sample_factor, centred.mT, upper=False).mT
        factor, info = torch.linalg.cholesky_ex(cov64)
        if bool((info != 0).any()):  # semi-definite target (a zero sigma): symmetric square root
            values, vectors = torch.linalg.eigh(cov64)
            factor = vectors * values.clamp_min(0.0).sqrt().unsqueeze(-2)
        vector_shape = torch.broadcast_shapes(mu64.shape[:-1], cov64.shape[:-2])
        samples = (
            standard @ factor.expand(*vector_shape, 6, 6).mT
            + mu64.expand(*vector_shape, 6).unsqueeze(-2)
        )
        particles = torch.cat([samples, torch.ones_like(samples[..., :1])], dim=-1)
        beam = cls(
            particles.to(**factory_kwargs),
            energy.to(**factory_kwargs),
            particle_charges=particle_charges,
            s=s,
            species=species,
            **factory_kwargs,
        )
        beam._unit_seventh = True
        return beam

    @classmethod
    def from_parameters(
        cls,
        num_particles: int = 100_000,
        mu_x=0.0, mu_px=0.0, mu_y=0.0, mu_py=0.0, mu_tau=0.0, mu_p=0.0,
        sigma_x=175e-6, sigma_px=4e-6, sigma_y=175e-6, sigma_py=4e-6,
        sigma_tau=8e-6, sigma_p=2e-3,
        cov_xpx=0.0, cov_ypy=0.0, cov_taup=0.0,
        energy: torch.Tensor | None = None,
        total_charge: torch.Tensor | None = None,
        s: torch.Tensor | None = None,
        species: Species | None = None,
        device: torch.device | None = None,
        dtype: torch.dtype | None = None,
        generator: torch.Generator | None = None,
        **other_covariances,
    ) -> "ParticleBeam":
        """Defaults follow particle_beam.py:199-216; numbers or tensors of mutually broadcastable
        shapes (a vectorised beam); the other twelve ``cov_*`` of particle_beam.py:109-140 are
        accepted by keyword."""
        t = lambda v: torch.as_tensor(v if v is not None else 0.0).detach().to(  # noqa: E731
            "cpu", torch.float64)
        means = [t(v) for v in (mu_x, mu_px, mu_y, mu_py, mu_tau, mu_p)]
        sigmas = [t(v) for v in (sigma_x, sigma_px, sigma_y, sigma_py, sigma_tau, sigma_p)]
        covariances = {"cov_xpx": t(cov_xpx), "cov_ypy": t(cov_ypy), "cov_taup": t(cov_taup)}
        for name, value in other_covariances.items():
            assert name in COVARIANCES, f"unknown beam parameter {name!r}"
            covariances[name] = t(value)
        mu = torch.stack(torch.broadcast_tensors(*means), dim=-1)
        shape = torch.broadcast_shapes(*[v.shape for v in sigmas + list(covariances.values())])
        cov = torch.zeros((*shape, 6, 6), dtype=torch.float64)
        for i, sigma in enumerate(sigmas):
            cov[..., i, i] = sigma.square()
        for name, value in covariances.items():
            i, j = COVARIANCES[name]
            cov[..., i, j] = cov[..., j, i] = value
        return cls.from_distribution(
            mu, cov, num_particles, energy, total_charge, s, species, device, dtype, generator
        )

    @classmethod
    def from_twiss(
        cls,
        num_particles: int = 100_000,
        beta_x=0.0, alpha_x=0.0, emittance_x=7.1971891e-13,
        beta_y=0.0, alpha_y=0.0, emittance_y=7.1971891e-13,
        sigma_tau=1e-6, sigma_p=1e-6, cov_taup=0.0,
        dispersion_x=0.0, dispersion_px=0.0, dispersion_y=0.0, dispersion_py=0.0,
        energy: torch.Tensor | None = None,
        total_charge: torch.Tensor | None = None,
        s: torch.Tensor | None = None,
        species: Species | None = None,
        device: torch.device | None = None,
        dtype: torch.dtype | None = None,
        generator: torch.Generator | None = None,
    ) -> "ParticleBeam":
        """Twiss and dispersion functions -> second moments as in particle_beam.py:434-561."""
        t = lambda v: torch.as_tensor(v).detach().to("cpu", torch.float64)  # noqa: E731
        beta_x, alpha_x, emittance_x = t(beta_x), t(alpha_x), t(emittance_x)
        beta_y, alpha_y, emittance_y = t(beta_y), t(alpha_y), t(emittance_y)
        sigma_p = t(sigma_p)
        dx, dpx, dy, dpy = t(dispersion_x), t(dispersion_px), t(dispersion_y), t(dispersion_py)
        assert (beta_x > 0).all(), "Beta function in x direction must be larger than 0 everywhere."
        assert (beta_y > 0).all(), "Beta function in y direction must be larger than 0 everywhere."
        p2 = sigma_p.square()
        return cls.from_parameters(
            num_particles,
            sigma_x=(emittance_x * beta_x + dx.square() * p2).sqrt(),
            sigma_px=(emittance_x * (1 + alpha_x.square()) / beta_x + dpx.square() * p2).sqrt(),
            sigma_y=(emittance_y * beta_y + dy.square() * p2).sqrt(),
            sigma_py=(emittance_y * (1 + alpha_y.square()) / beta_y + dpy.square() * p2).sqrt(),
            sigma_tau=sigma_tau, sigma_p=sigma_p,
            cov_xpx=-emittance_x * alpha_x + dx * dpx * p2,
            cov_ypy=-emittance_y * alpha_y + dy * dpy * p2,
            cov_taup=cov_taup,
            cov_xp=dx * p2, cov_pxp=dpx * p2, cov_yp=dy * p2, cov_pyp=dpy * p2,
            energy=energy, total_charge=total_charge, s=s, species=species,
            device=device, dtype=dtype, generator=generator,
        )

    @classmethod
    def uniform_3d_ellipsoid(cls, num_particles: int = 100_000, radius_x=1e-3, radius_y=1e-3,
                             radius_tau=1e-3, sigma_px=4e-6, sigma_py=4e-6, sigma_p=2e-3,
                             energy=None, total_charge=None, s=None, species=None, device=None,
                             dtype=None, generator: torch.Generator | None = None
                             ) -> "ParticleBeam":
        """Waterbag beam: positions uniform inside an ellipsoid, Gaussian uncorrelated momenta
        (particle_beam.py:563-666; the space-charge expansion test of the reference starts from
        it)."""
        beam = cls.from_parameters(
            num_particles, sigma_x=radius_x, sigma_px=sigma_px, sigma_y=radius_y,
            sigma_py=sigma_py, sigma_tau=radius_tau, sigma_p=sigma_p, energy=energy,
            total_charge=total_charge, s=s, species=species, device=device, dtype=dtype,
            generator=generator,
        )
        uniform = torch.rand(3, num_particles, dtype=torch.float64, generator=generator)
        r = uniform[0].pow(1 / 3)                   # uniform in the volume of the unit sphere
        theta = (2 * uniform[1] - 1).arccos()
        phi = uniform[2] * 2 * torch.pi
        unit = torch.stack([r * theta.sin() * phi.cos(), r * theta.sin() * phi.sin(),
                            r * theta.cos()])
        radii = [float(radius_x), float(radius_y), float(radius_tau)]
        for column, radius, values in zip((0, 2, 4), radii, unit):
            beam.particles[..., column] = (values * radius).to(beam.particles)
        return beam

    @classmethod
    def make_linspaced(cls, num_particles: int = 10, energy=None, total_charge=None, s=None,
                       species=None, device=None, dtype=None, **moments) -> "ParticleBeam":
        """``num_particles`` particles evenly spaced from ``mu - sigma`` to ``mu + sigma`` in every
        coordinate; ``mu_*`` / ``sigma_*`` by keyword, tensors of broadcastable shapes
        (particle_beam.py:668-803)."""
        factory_kwargs = {"device": device, "dtype": dtype}
        defaults = {"sigma_x": 175e-9, "sigma_px": 2e-7, "sigma_y": 175e-9, "sigma_py": 2e-7,
                    "sigma_tau": 1e-6, "sigma_p": 1e-6}
        for name in moments:
            assert name in {f"{k}_{c}" for k in ("mu", "sigma") for c in COORDINATES}, (
                f"unknown beam parameter {name!r}"
            )
        value = lambda name: torch.as_tensor(  # noqa: E731
            moments[name] if moments.get(name) is not None else defaults.get(name, 0.0),
            **factory_kwargs)
        species = species if species is not None else Species("electron", **factory_kwargs)
        energy = energy if energy is not None else torch.tensor(1e8, **factory_kwargs)
        total_charge = torch.as_tensor(
            total_charge if total_charge is not None else species.charge_coulomb * num_particles,
            **factory_kwargs)
        charges = (torch.ones((*total_charge.shape, num_particles), **factory_kwargs)
                   * total_charge.unsqueeze(-1) / num_particles)
        centres = [value(f"mu_{c}") for c in COORDINATES]
        widths = [value(f"sigma_{c}") for c in COORDINATES]
        vector_shape = torch.broadcast_shapes(*[t.shape for t in centres + widths])
        particles = torch.ones((*vector_shape, num_particles, 7), **factory_kwargs)
        steps = torch.linspace(-1.0, 1.0, num_particles, **factory_kwargs)
        for i, (mu, sigma) in enumerate(zip(centres, widths)):
            particles[..., i] = mu.unsqueeze(-1) + sigma.unsqueeze(-1) * steps
        return cls(particles, energy, particle_charges=charges, s=s, species=species,
                   device=device, dtype=dtype)

    def linspaced(self, num_particles: int) -> "ParticleBeam":
        """Evenly spaced beam with this beam's centres, sigmas, energy and total charge
        (particle_beam.py:1180-1210)."""
        moments = {f"{k}_{c}": getattr(self, f"{k}_{c}") for k in ("mu", "sigma")
                   for c in COORDINATES}
        return self.make_linspaced(
            num_particles, energy=self.energy, total_charge=self.total_charge, s=self.s,
            species=self.species, device=self.particles.device, dtype=self.particles.dtype,
            **moments,
        )

    # ---- views and moments -----------------------------------------------------------------
    @property
    def num_particles(self) -> int:
        return self.particles.shape[-2]

    def __len__(self) -> int:
        return int(self.num_particles)

    @property
    def total_charge(self) -> torch.Tensor:
        return (self.particle_charges * self.survival_probabilities).sum(dim=-1)

    @property
    def num_particles_survived(self) -> torch.Tensor:
        return self.survival_probabilities.sum(dim=-1)

    def _coordinate(self, index: int) -> torch.Tensor:
        return self.particles[..., index]

    def _mean(self, index: int) -> torch.Tensor:
        w = self.survival_probabilities
        return (self._coordinate(index) * w).sum(dim=-1) / w.sum(dim=-1)

    def _std(self, index: int) -> torch.Tensor:
        # unbiased survival-weighted std: cheetah/utils/statistics.py:30-62
        w = self.survival_probabilities
        v = self._coordinate(index)
        sum_w = w.sum(dim=-1)
        mean = (v * w).sum(dim=-1) / sum_w
        correction = sum_w - w.square().sum(dim=-1) / sum_w
        return ((w * (v - mean.unsqueeze(-1)).square()).sum(dim=-1) / correction).sqrt()

    def second_moments(self) -> tuple[torch.Tensor, torch.Tensor]:
        """Survival-weighted means ``(..., 6)`` and unbiased covariance ``(..., 6, 6)`` of the six
        coordinates (cheetah/utils/statistics.py:65-88; what ``as_parameter_beam``, the ``cov_*``
        properties, emittances and Twiss functions are made of).  On a CUDA device this is one
        pass of the fused covariance kernel (``ch_apply_maps_covariance`` with the identity map:
        fp64 sums about a pilot particle, no centred copy of the beam); CPU beams -- set-up only --
        use the textbook formula."""
        if self.particles.is_cuda:
            from . import tracking

            observed = tracking.beam_moments(self)
            return observed.mu, observed.cov
        w = self.survival_probabilities.unsqueeze(-1)
        u = self.particles[..., :6]
        sum_w = w.sum(dim=-2, keepdim=True)
        mean = (u * w).sum(dim=-2, keepdim=True) / sum_w
        centred = u - mean
        correction = sum_w - w.square().sum(dim=-2, keepdim=True) / sum_w
        return mean.squeeze(-2), (w * centred).mT @ centred / correction

    def _covariance(self, i: int, j: int) -> torch.Tensor:
        return self.second_moments()[1][..., i, j]

    def as_parameter_beam(self) -> "ParameterBeam":
        """ParameterBeam with this beam's moments (particle_beam.py:1160-1178)."""
        mean, covariance = self.second_moments()
        mu = torch.cat([mean, torch.ones_like(mean[..., :1])], dim=-1)
        cov = torch.zeros((*covariance.shape[:-2], 7, 7), dtype=covariance.dtype,
                          device=covariance.device)
        cov[..., :6, :6] = covariance
        return ParameterBeam(mu, cov, self.energy, total_charge=self.total_charge,
                             device=mu.device, dtype=mu.dtype)

    def transformed_to(self, energy: torch.Tensor | None = None,
                       total_charge: torch.Tensor | None = None,
                       species: Species | None = None, **moments) -> "ParticleBeam":
        """Shift and rescale every coordinate to new ``mu_*`` / ``sigma_*`` (omitted ones keep
        their value); like the reference (particle_beam.py:1034-1158) the covariance arguments
        of the ParameterBeam signature are not applicable to a particle cloud."""
        for name in moments:
            assert name in {f"{k}_{c}" for k in ("mu", "sigma") for c in COORDINATES}, (
                f"ParticleBeam.transformed_to cannot set {name!r}"
            )
        old_mu = torch.stack([getattr(self, f"mu_{c}") for c in COORDINATES], dim=-1)
        old_sigma = torch.stack([getattr(self, f"sigma_{c}") for c in COORDINATES], dim=-1)
        pick = lambda kind, c, old: torch.as_tensor(  # noqa: E731
            moments.get(f"{kind}_{c}", old), dtype=old.dtype, device=old.device)
        new_mu = torch.stack(torch.broadcast_tensors(
            *[pick("mu", c, old_mu[..., i]) for i, c in enumerate(COORDINATES)]), dim=-1)
        new_sigma = torch.stack(torch.broadcast_tensors(
            *[pick("sigma", c, old_sigma[..., i]) for i, c in enumerate(COORDINATES)]), dim=-1)
        phase_space = (
            (self.particles[..., :6] - old_mu.unsqueeze(-2)) / old_sigma.unsqueeze(-2)
            * new_sigma.unsqueeze(-2) + new_mu.unsqueeze(-2)
        )
        particles = torch.cat([phase_space, torch.ones_like(phase_space[..., :1])], dim=-1)
        if total_charge is None:
            particle_charges = self.particle_charges
        elif self.total_charge is None:  # scale to the new charge
            total_charge = torch.as_tensor(total_charge, dtype=particles.dtype,
                                           device=particles.device)
            particle_charges = self.particle_charges * (
                total_charge / self.total_charge).unsqueeze(-1)
        else:  # (the reference's branch for every charged beam: spread the charge evenly)
            total_charge = torch.as_tensor(total_charge, dtype=particles.dtype,
                                           device=particles.device)
            particle_charges = (
                torch.ones_like(self.particle_charges, device=total_charge.device)
                * total_charge.unsqueeze(-1) / self.particle_charges.shape[-1]
            )
        return self.__class__(
            particles, energy if energy is not None else self.energy,
            particle_charges=particle_charges,
            survival_probabilities=self.survival_probabilities, s=self.s,
            species=species if species is not None else self.species,
        )

    @property
    def energies(self) -> torch.Tensor:
        """Energies of the individual particles in eV (particle_beam.py:1945-1948)."""
        return self.p * self.p0c.unsqueeze(-1) + self.energy.unsqueeze(-1)

    @property
    def momenta(self) -> torch.Tensor:
        """Momenta of the individual particles in eV/c (particle_beam.py:1950-1953)."""
        return (self.energies.square() - self.species.mass_eV.square()).sqrt()

    def to_xyz_pxpypz(self) -> torch.Tensor:
        """``(x, Px, y, Py, z, Pz, 1)`` in SI units (particle_beam.py:1316-1346)."""
        gamma0 = self.relativistic_gamma.unsqueeze(-1)
        beta0 = self.relativistic_beta.unsqueeze(-1)
        mc = self.species.mass_kg * SPEED_OF_LIGHT
        p0 = gamma0 * beta0 * mc
        gamma = gamma0 * (1.0 + self.particles[..., 5] * beta0)
        momentum = gamma * mc * (1 - gamma.square().reciprocal()).sqrt()
        out = self.particles.clone()
        out[..., 1] = self.particles[..., 1] * p0
        out[..., 3] = self.particles[..., 3] * p0
        out[..., 4] = -self.particles[..., 4] * beta0
        out[..., 5] = (momentum.square() - out[..., 1].square() - out[..., 3].square()).sqrt()
        return out

    @classmethod
    def from_xyz_pxpypz(cls, xp_coordinates: torch.Tensor, energy: torch.Tensor,
                        particle_charges: torch.Tensor | None = None,
                        survival_probabilities: torch.Tensor | None = None,
                        s: torch.Tensor | None = None, species: Species | None = None,
                        device: torch.device | None = None,
                        dtype: torch.dtype | None = None) -> "ParticleBeam":
        """Inverse of ``to_xyz_pxpypz`` (particle_beam.py:1262-1314)."""
        beam = cls(xp_coordinates.clone(), energy, particle_charges, survival_probabilities, s,
                   species, device, dtype)
        gamma0 = beam.relativistic_gamma.unsqueeze(-1)
        beta0 = beam.relativistic_beta.unsqueeze(-1)
        mc = beam.species.mass_kg * SPEED_OF_LIGHT
        p0 = gamma0 * beta0 * mc
        p = (xp_coordinates[..., 1].square() + xp_coordinates[..., 3].square()
             + xp_coordinates[..., 5].square()).sqrt()
        gamma = (1 + (p / mc).square()).sqrt()
        beam.particles[..., 1] = xp_coordinates[..., 1] / p0
        beam.particles[..., 3] = xp_coordinates[..., 3] / p0
        beam.particles[..., 4] = -xp_coordinates[..., 4] / beta0
        beam.particles[..., 5] = (gamma - gamma0) / (beta0 * gamma0)
        return beam

    def __getitem__(self, item) -> "ParticleBeam":
        """Index the vector dimensions (particle_beam.py:1976-2001)."""
        vector_shape = torch.broadcast_shapes(
            self.particles.shape[:-2], self.energy.shape, self.particle_charges.shape[:-1],
            self.survival_probabilities.shape[:-1],
        )
        n = self.num_particles
        return self.__class__(
            particles=self.particles.broadcast_to((*vector_shape, n, 7))[item],
            energy=self.energy.broadcast_to(vector_shape)[item],
            particle_charges=self.particle_charges.broadcast_to((*vector_shape, n))[item],
            survival_probabilities=self.survival_probabilities.broadcast_to(
                (*vector_shape, n))[item],
            device=self.particles.device, dtype=self.particles.dtype,
        )

    def randomly_subsampled(self, num_particles: int, adjust_particle_charges: bool = True,
                            random_state: torch.Generator | None = None) -> "ParticleBeam":
        """Beam of ``num_particles`` particles drawn without replacement; with
        ``adjust_particle_charges`` their charges are rescaled to the old total charge
        (particle_beam.py:1212-1260)."""
        assert num_particles <= self.num_particles, (
            "Number of particles to sample must be less than or equal to the number of "
            "particles in the original beam."
        )
        keep = torch.randperm(self.num_particles, generator=random_state,
                              device=self.particles.device)[:num_particles]
        beam = self.__class__(
            self.particles[..., keep, :], self.energy,
            particle_charges=self.particle_charges[..., keep],
            survival_probabilities=self.survival_probabilities[..., keep],
            species=self.species,
        )
        if adjust_particle_charges:
            beam.particle_charges = beam.particle_charges * (
                self.total_charge / beam.total_charge).unsqueeze(-1)
        return beam

    def clone(self) -> "ParticleBeam":
        beam = self.__class__(
            particles=self.particles.clone(),
            energy=self.energy.clone(),
            particle_charges=self.particle_charges.clone(),
            survival_probabilities=self.survival_probabilities.clone(),
            s=self.s.clone(),
            species=self.species.clone(),
        )
        beam._unit_seventh = self._unit_seventh
        return beam

    def __repr__(self) -> str:
        return (
            f"{self.__class__.__name__}(particles={tuple(self.particles.shape)}, "
            f"energy={self.energy!r}, s={self.s!r}, species={self.species!r})"
        )


for _i, _name in enumerate(("x", "px", "y", "py", "tau", "p")):
    setattr(ParticleBeam, _name, property(lambda self, i=_i: self._coordinate(i)))
    setattr(ParticleBeam, f"mu_{_name}", property(lambda self, i=_i: self._mean(i)))
    setattr(ParticleBeam, f"sigma_{_name}", property(lambda self, i=_i: self._std(i)))


class ParameterBeam(Beam):
    """Gaussian-moment beam: ``mu (..., 7)``, ``cov (..., 7, 7)`` (parameter_beam.py:8-60)."""

    def __init__(
        self,
        mu: torch.Tensor,
        cov: torch.Tensor,
        energy: torch.Tensor,
        total_charge: torch.Tensor | None = None,
        s: torch.Tensor | None = None,
        species: Species | None = None,
        device: torch.device | None = None,
        dtype: torch.dtype | None = None,
    ) -> None:
        super().__init__()
        device = device if device is not None else mu.device
        dtype = dtype if dtype is not None else mu.dtype
        factory_kwargs = {"device": device, "dtype": dtype}
        self.species = species if species is not None else Species("electron", **factory_kwargs)
        self.register_buffer("mu", mu)
        self.register_buffer("cov", cov)
        self.register_buffer("energy", energy)
        self.register_buffer(
            "total_charge",
            total_charge if total_charge is not None else torch.tensor(0.0, **factory_kwargs),
        )
        self.register_buffer("s", s if s is not None else torch.tensor(0.0, **factory_kwargs))

    @classmethod
    def from_parameters(cls, energy: torch.Tensor | None = None,
                        total_charge: torch.Tensor | None = None, s: torch.Tensor | None = None,
                        species: Species | None = None, device: torch.device | None = None,
                        dtype: torch.dtype | None = None, **moments) -> "ParameterBeam":
        """``mu_*``, ``sigma_*`` and the fifteen ``cov_*`` by keyword, tensors of any mutually
        broadcastable shapes; defaults and the positive-definiteness check follow
        parameter_beam.py:62-281."""
        factory_kwargs = {"device": device, "dtype": dtype}
        defaults = {"sigma_x": 175e-6, "sigma_px": 4e-6, "sigma_y": 175e-6, "sigma_py": 4e-6,
                    "sigma_tau": 8e-6, "sigma_p": 2e-3}
        allowed = {f"mu_{c}" for c in COORDINATES} | set(defaults) | set(COVARIANCES)
        for name in moments:
            assert name in allowed, f"unknown beam parameter {name!r}"
        value = lambda name: torch.as_tensor(  # noqa: E731
            moments[name] if moments.get(name) is not None else defaults.get(name, 0.0),
            **factory_kwargs)
        means = torch.broadcast_tensors(*[value(f"mu_{c}") for c in COORDINATES])
        mu = torch.stack([*means, torch.ones_like(means[0])], dim=-1)
        names = [f"sigma_{c}" for c in COORDINATES] + list(COVARIANCES)
        entries = dict(zip(names, torch.broadcast_tensors(*[value(n) for n in names])))
        cov = torch.zeros(*entries["sigma_x"].shape, 7, 7, **factory_kwargs)
        for i, c in enumerate(COORDINATES):
            cov[..., i, i] = entries[f"sigma_{c}"].square()
        for name, (i, j) in COVARIANCES.items():
            cov[..., i, j] = cov[..., j, i] = entries[name]
        try:
            torch.linalg.cholesky(cov[..., :6, :6])
        except RuntimeError as e:
            raise ValueError(
                "The covariance matrix of the beam must be positive definite. Please check the "
                "input parameters to ensure that they are consistent."
            ) from e
        return cls(
            mu, cov,
            energy if energy is not None else torch.tensor(1e8, **factory_kwargs),
            total_charge=total_charge, s=s, species=species, device=device, dtype=dtype,
        )

    @classmethod
    def from_twiss(cls, beta_x=None, alpha_x=None, emittance_x=None, beta_y=None, alpha_y=None,
                   emittance_y=None, sigma_tau=None, sigma_p=None, cov_taup=None,
                   dispersion_x=None, dispersion_px=None, dispersion_y=None, dispersion_py=None,
                   energy=None, total_charge=None, s=None, species=None, device=None,
                   dtype=None) -> "ParameterBeam":
        """Twiss and dispersion functions -> second moments (parameter_beam.py:283-414)."""
        factory_kwargs = {"device": device, "dtype": dtype}
        t = lambda v, default: torch.as_tensor(  # noqa: E731
            v if v is not None else default, **factory_kwargs)
        beta_x, alpha_x, emittance_x = t(beta_x, 0.0), t(alpha_x, 0.0), t(emittance_x, 7.1971891e-13)
        beta_y, alpha_y, emittance_y = t(beta_y, 0.0), t(alpha_y, 0.0), t(emittance_y, 7.1971891e-13)
        sigma_tau, sigma_p, cov_taup = t(sigma_tau, 1e-6), t(sigma_p, 1e-6), t(cov_taup, 0.0)
        dx, dpx, dy, dpy = (t(d, 0.0) for d in (dispersion_x, dispersion_px, dispersion_y,
                                                dispersion_py))
        assert (beta_x > 0).all(), "Beta function in x direction must be larger than 0 everywhere."
        assert (beta_y > 0).all(), "Beta function in y direction must be larger than 0 everywhere."
        p2 = sigma_p.square()
        return cls.from_parameters(
            sigma_x=(emittance_x * beta_x + dx.square() * p2).sqrt(),
            sigma_px=(emittance_x * (1 + alpha_x.square()) / beta_x + dpx.square() * p2).sqrt(),
            sigma_y=(emittance_y * beta_y + dy.square() * p2).sqrt(),
            sigma_py=(emittance_y * (1 + alpha_y.square()) / beta_y + dpy.square() * p2).sqrt(),
            sigma_tau=sigma_tau, sigma_p=sigma_p, cov_taup=cov_taup,
            cov_xpx=-emittance_x * alpha_x + dx * dpx * p2,
            cov_ypy=-emittance_y * alpha_y + dy * dpy * p2,
            cov_xp=dx * p2, cov_pxp=dpx * p2, cov_yp=dy * p2, cov_pyp=dpy * p2,
            energy=energy, total_charge=total_charge, s=s, species=species, device=device,
            dtype=dtype,
        )

    def transformed_to(self, energy=None, total_charge=None, species=None,
                       **moments) -> "ParameterBeam":
        """New beam with the given parameters replaced (parameter_beam.py:476-586)."""
        current = {f"mu_{c}": getattr(self, f"mu_{c}") for c in COORDINATES}
        current.update({f"sigma_{c}": getattr(self, f"sigma_{c}") for c in COORDINATES})
        current.update({name: getattr(self, name) for name in COVARIANCES})
        current.update({k: v for k, v in moments.items() if v is not None})
        return self.__class__.from_parameters(
            energy=energy if energy is not None else self.energy,
            total_charge=total_charge if total_charge is not None else self.total_charge,
            s=self.s, species=species if species is not None else self.species,
            device=self.mu.device, dtype=self.mu.dtype, **current,
        )

    def as_particle_beam(self, num_particles: int,
                         generator: torch.Generator | None = None) -> ParticleBeam:
        """Random ParticleBeam with this beam's moments (parameter_beam.py:588-608;
        non-vectorised)."""
        assert self.mu.dim() == 1, "as_particle_beam needs a non-vectorised ParameterBeam"
        return ParticleBeam.from_distribution(
            self.mu[:6], self.cov[:6, :6], num_particles, self.energy, self.total_charge, self.s,
            self.species, self.mu.device, self.mu.dtype, generator,
        )

    def linspaced(self, num_particles: int) -> ParticleBeam:
        """ParticleBeam of evenly spaced particles with this beam's centres and sigmas
        (parameter_beam.py:610-640)."""
        moments = {f"{k}_{c}": getattr(self, f"{k}_{c}") for k in ("mu", "sigma")
                   for c in COORDINATES}
        return ParticleBeam.make_linspaced(
            num_particles, energy=self.energy, total_charge=self.total_charge, s=self.s,
            species=self.species, device=self.mu.device, dtype=self.mu.dtype, **moments,
        )

    def _covariance(self, i: int, j: int) -> torch.Tensor:
        return self.cov[..., i, j]

    def clone(self) -> "ParameterBeam":
        return self.__class__(
            self.mu.clone(), self.cov.clone(), self.energy.clone(),
            total_charge=self.total_charge.clone(), s=self.s.clone(),
            species=self.species.clone(),
        )

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}(mu={self.mu!r}, energy={self.energy!r}, s={self.s!r})"


for _i, _name in enumerate(COORDINATES):
    setattr(ParameterBeam, f"mu_{_name}", property(lambda self, i=_i: self.mu[..., i]))
    setattr(ParameterBeam, f"sigma_{_name}",
            property(lambda self, i=_i: self.cov[..., i, i].sqrt()))
