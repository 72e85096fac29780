"""Data boundary of the hot path: GPU-resident mirror of the reference's data-set classes (diff_gfdn/dataloader.py).

The reference builds its frequency-domain targets with numpy on the host (`RoomDataset.__init__` :188-254,
`early_late_split` :300-325), wraps them in a torch `Dataset` whose `__getitem__` returns one receiver at a time
(:578-600) and stacks every batch in Python (`custom_collate` :674-704). Here the same arrays are produced ON the
device (cuFFT through torch.fft, float64 like numpy, stored complex64 -- what the kernels consume), stay resident,
and a batch is one `index_select` per field. Class names, constructor arguments, attribute names, the batch
dictionary keys and the split semantics (`create_fixed_test_split` :707-724 with its seeded `randperm`,
`split_dataset` :727-745) are the reference's. File IO (pickle / wav / SOFA readers, `ThreeRoomDataset`) is out of
scope: construct from arrays.
"""
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Union

import numpy as np
import torch
from torch.utils import data

from .utils import ms_to_samps

C64 = torch.complex64


@dataclass
class Meshgrid:
    xmesh: torch.Tensor
    ymesh: torch.Tensor


def _need_cuda(device) -> torch.device:
    dev = torch.device(device if device is not None else "cuda")
    if dev.type != "cuda":
        raise RuntimeError("diffgfdn_b200.dataloader: data sets are GPU resident, pass a CUDA device (no CPU path)")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


def _early_late_split(rirs: torch.Tensor, sample_rate: float, mixing_time_ms: float, nfft: int, win_len_ms: float = 5.0):
    """reference :300-325 / :156-182: hann cross-fade around the mixing time, applied IN PLACE to `rirs` (the
    reference's slices are views of its array too), then rfft of both parts. Returns (early, late) complex64."""
    mix = ms_to_samps(mixing_time_ms, sample_rate)
    wl = ms_to_samps(win_len_ms, sample_rate)
    window = torch.tensor(np.hanning(wl), dtype=rirs.dtype, device=rirs.device)
    fade_in, fade_out = window[:wl // 2], window[wl // 2:]
    early, late = rirs[..., :mix], rirs[..., mix:]
    early[..., -fade_out.numel():] *= fade_out  # the reference's [-win_len // 2:] is floor(-wl / 2) samples
    late[..., :wl // 2] *= fade_in
    return (torch.fft.rfft(early, n=nfft, dim=-1).to(C64), torch.fft.rfft(late, n=nfft, dim=-1).to(C64))


class RIRData:
    """One measured / simulated RIR (reference :76-182). `rir_mag_response` is taken AFTER the early/late fades have
    been written into the RIR, as in the reference (there it is a property evaluated lazily)."""

    def __init__(self, common_decay_times, band_centre_hz=None, amplitudes=None, room_dims=None, absorption_coeffs=None,
                 mixing_time_ms: float = 20.0, nfft: Optional[int] = None, wav_path=None, rir=None,
                 sample_rate: Optional[float] = None, device=None):
        if rir is None:
            raise AttributeError("RIRData: pass the RIR as an array (reading .wav files is outside this package)")
        self.device = _need_cuda(device)
        self.sample_rate = sample_rate
        self.common_decay_times = common_decay_times
        self.band_centre_hz = band_centre_hz
        self.amplitudes = amplitudes
        self.mixing_time_ms = mixing_time_ms
        self.room_dims = room_dims
        self.absorption_coeffs = absorption_coeffs
        self.nfft = nfft
        self.rir = torch.as_tensor(np.asarray(rir), dtype=torch.float64).to(self.device).clone()
        self.early_rir_mag_response, self.late_rir_mag_response = _early_late_split(
            self.rir, self.sample_rate, self.mixing_time_ms, self.num_freq_bins)
        self.rir_mag_response = torch.fft.rfft(self.rir, n=self.num_freq_bins).to(C64)

    @property
    def num_freq_bins(self) -> int:
        if self.nfft is not None:
            return self.nfft
        return int(np.power(2, np.ceil(np.log2(np.max(self.common_decay_times) * self.sample_rate))))

    @property
    def freq_bins_rad(self):
        return np.fft.rfftfreq(self.num_freq_bins) * 2 * np.pi

    @property
    def freq_bins_hz(self):
        return np.fft.rfftfreq(self.num_freq_bins, d=1.0 / self.sample_rate)


class RoomDataset:
    """RIRs of one space over a grid of receivers (and sources), reference :185-423, built from arrays."""

    def __init__(self, num_rooms: int, sample_rate: float, source_position, receiver_position, rirs, common_decay_times,
                 room_dims: List, room_start_coord: List, band_centre_hz=None, amplitudes=None, noise_floor=None,
                 absorption_coeffs=None, aperture_coords=None, mixing_time_ms: float = 20.0, nfft: Optional[int] = None,
                 grid_spacing_m: float = 0.3, device=None):
        self.device = _need_cuda(device)
        self.sample_rate = sample_rate
        self.num_rooms = num_rooms
        self.source_position = np.asarray(source_position)
        self.receiver_position = np.asarray(receiver_position)
        self.band_centre_hz = band_centre_hz
        self.common_decay_times = np.asarray(common_decay_times)
        self.noise_floor = noise_floor
        self.amplitudes = amplitudes
        self.num_rec = self.receiver_position.shape[0]
        self.num_src = self.source_position.shape[0] if self.source_position.ndim > 1 else 1
        self.absorption_coeffs = absorption_coeffs
        self.room_dims = room_dims
        self.room_start_coord = room_start_coord
        self.aperture_coords = aperture_coords
        self.mixing_time_ms = mixing_time_ms
        self.nfft = nfft
        self._eps = 1e-12
        self.rirs = torch.as_tensor(np.asarray(rirs), dtype=torch.float64).to(self.device).clone()
        self.rir_length = self.rirs.shape[-1]
        # full response BEFORE the fades are written into the RIRs (reference :252-253)
        self.rir_mag_response = torch.fft.rfft(self.rirs, n=self.num_freq_bins, dim=-1).to(C64)
        self.early_late_split()
        self.grid_spacing_m = grid_spacing_m
        self.mesh_2D = self.get_2D_meshgrid()

    @property
    def norm_receiver_position(self) -> np.ndarray:
        p = self.receiver_position
        lo, hi = p.min(axis=0), p.max(axis=0)
        return (p - lo) / ((hi - lo) + self._eps)

    @property
    def num_freq_bins(self) -> int:
        if self.nfft is not None:
            return self.nfft
        return int(np.power(2, np.ceil(np.log2(self.common_decay_times.max() * self.sample_rate))))

    @property
    def freq_bins_rad(self):
        return np.fft.rfftfreq(self.num_freq_bins) * 2 * np.pi

    @property
    def freq_bins_hz(self):
        return np.fft.rfftfreq(self.num_freq_bins, d=1.0 / self.sample_rate)

    def find_rec_idx_in_room_dataset(self, rec_pos_list) -> np.ndarray:
        dist = np.linalg.norm(self.receiver_position[:, None, :] - np.asarray(rec_pos_list), axis=2)
        return np.argmin(dist, axis=0)

    def early_late_split(self, win_len_ms: float = 5.0):
        self.early_rir_mag_response, self.late_rir_mag_response = _early_late_split(
            self.rirs, self.sample_rate, self.mixing_time_ms, self.num_freq_bins, win_len_ms)

    def update_receiver_pos(self, new_receiver_pos):
        self.receiver_position = np.asarray(new_receiver_pos)
        self.num_rec = self.receiver_position.shape[0]

    def get_2D_meshgrid(self) -> Meshgrid:
        xs, ys = [], []
        for nroom in range(self.num_rooms):
            nx = int(self.room_dims[nroom][0] / self.grid_spacing_m)
            ny = int(self.room_dims[nroom][1] / self.grid_spacing_m)
            x = np.linspace(self.room_start_coord[nroom][0], self.room_start_coord[nroom][0] + self.room_dims[nroom][0], nx)
            y = np.linspace(self.room_start_coord[nroom][1], self.room_start_coord[nroom][1] + self.room_dims[nroom][1], ny)
            xm, ym = np.meshgrid(x, y)
            xs.append(xm.flatten())
            ys.append(ym.flatten())
        return Meshgrid(torch.from_numpy(np.concatenate(xs)), torch.from_numpy(np.concatenate(ys)))


def _z_values(freq_bins_rad, new_sampling_radius, device) -> torch.Tensor:
    w = torch.as_tensor(freq_bins_rad, dtype=torch.float64, device=device)
    r = 1.0 if new_sampling_radius in (1.0, None) else float(new_sampling_radius)
    assert r >= 1.0
    return torch.polar(torch.full_like(w, r), w)


class MultiRIRDataset(data.Dataset):
    """Receiver grid data set (reference :515-600). Item idx is one (source, receiver) pair; `batch(indices)` returns
    the collated dictionary of `custom_collate` for a whole index tensor with one gather per field."""

    def __init__(self, device, room_data: RoomDataset, new_sampling_radius: Optional[float] = None):
        self.device = _need_cuda(device)
        dev = self.device
        src = torch.as_tensor(room_data.source_position, device=dev)
        self.source_position = src.unsqueeze(0) if src.dim() == 1 else src
        self.listener_positions = torch.as_tensor(room_data.receiver_position, device=dev)
        self.norm_listener_position = torch.as_tensor(room_data.norm_receiver_position, device=dev)
        self.num_src = self.source_position.shape[0]
        self.num_rec = self.listener_positions.shape[0]
        self.index_pairs = [(i, j) for i in range(self.num_src) for j in range(self.num_rec)]
        self.z_values = _z_values(room_data.freq_bins_rad, new_sampling_radius, dev)
        self.rir_mag_response = room_data.rir_mag_response.to(dev)
        self.late_rir_mag_response = room_data.late_rir_mag_response.to(dev)
        self.early_rir_mag_response = room_data.early_rir_mag_response.to(dev)

    def __len__(self) -> int:
        return self.num_src * self.num_rec

    def _flat(self, t: torch.Tensor) -> torch.Tensor:
        return t if self.num_src == 1 and t.dim() == 2 else t.reshape(-1, t.shape[-1])

    def batch(self, indices: torch.Tensor) -> Dict:
        idx = torch.as_tensor(indices, device=self.device, dtype=torch.long).reshape(-1)
        i_src, i_rec = idx // self.num_rec, idx % self.num_rec
        return {
            'z_values': self.z_values,
            'source_position': self.source_position.index_select(0, i_src),
            'listener_position': self.listener_positions.index_select(0, i_rec),
            'norm_listener_position': self.norm_listener_position.index_select(0, i_rec),
            'target_early_response': self._flat(self.early_rir_mag_response).index_select(0, idx),
            'target_late_response': self._flat(self.late_rir_mag_response).index_select(0, idx),
            'target_rir_response': self._flat(self.rir_mag_response).index_select(0, idx),
        }

    def __getitem__(self, idx: int) -> Dict:
        b = self.batch(torch.tensor([int(idx)]))
        return {k: (v if k == 'z_values' else v[0]) for k, v in b.items()}


class SingleRIRDataset(data.Dataset):
    """One RIR; items are frequency bins (reference :603-658)."""

    def __init__(self, device, rir_data: RIRData, new_sampling_radius: Optional[float] = None):
        self.device = _need_cuda(device)
        self.z_values = _z_values(rir_data.freq_bins_rad, new_sampling_radius, self.device)
        self.rir_mag_response = rir_data.rir_mag_response.to(self.device)
        self.late_rir_mag_response = rir_data.late_rir_mag_response.to(self.device)
        self.early_rir_mag_response = rir_data.early_rir_mag_response.to(self.device)

    def __len__(self) -> int:
        return len(self.z_values)

    def batch(self, indices: torch.Tensor) -> Dict:
        idx = torch.as_tensor(indices, device=self.device, dtype=torch.long).reshape(-1)
        return {'z_values': self.z_values.index_select(0, idx),
                'target_rir_response': self.rir_mag_response.index_select(0, idx),
                'target_early_response': self.early_rir_mag_response.index_select(0, idx),
                'target_late_response': self.late_rir_mag_response.index_select(0, idx)}

    def __getitem__(self, idx: int) -> Dict:
        return {k: v[0] for k, v in self.batch(torch.tensor([int(idx)])).items()}


def custom_collate(batch: Sequence[Dict]) -> Dict:
    """Stack a list of items (API parity with reference :674-704; the loaders below never go through it)."""
    out = {'z_values': batch[0]['z_values']}
    for key in ('source_position', 'listener_position', 'norm_listener_position', 'target_early_response',
                'target_late_response', 'target_rir_response'):
        out[key] = torch.stack([item[key] for item in batch])
    return out


def to_device(data_class, device):
    """The data sets are built on their device; kept for API parity (reference :661-671)."""
    for name, value in list(data_class.__dict__.items()):
        if isinstance(value, torch.Tensor):
            setattr(data_class, name, value.to(device))
    return data_class


def _resolve(dataset):
    """(base data set, index tensor) of a data set or of nested torch Subsets."""
    idx = None
    while isinstance(dataset, data.Subset):
        cur = torch.as_tensor(dataset.indices, dtype=torch.long)
        idx = cur if idx is None else cur[idx]
        dataset = dataset.dataset
    if idx is None:
        idx = torch.arange(len(dataset))
    return dataset, idx


class GPUBatchLoader:
    """Iterable of collated batches of a GPU-resident data set: one randperm + one gather per field and batch instead
    of the reference's per-item `__getitem__` + Python `stack` (DataLoader + custom_collate, reference :748-772)."""

    def __init__(self, dataset, batch_size: int, shuffle: bool = True, drop_last: bool = True,
                 generator: Optional[torch.Generator] = None):
        self.dataset, idx = _resolve(dataset)
        self.indices = idx.to(self.dataset.device)
        self.batch_size = int(batch_size)
        self.shuffle = shuffle
        self.drop_last = drop_last
        self.generator = generator

    def __len__(self) -> int:
        n = self.indices.numel()
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        idx = self.indices
        if self.shuffle:
            perm = torch.randperm(idx.numel(), generator=self.generator, device='cpu').to(idx.device)
            idx = idx[perm]
        for b in range(len(self)):
            yield self.dataset.batch(idx[b * self.batch_size:(b + 1) * self.batch_size])


def create_fixed_test_split(dataset, test_ratio: float = 0.1, seed: int = 42):
    """reference :707-724: the same seeded CPU randperm, so the held-out receivers are the reference's."""
    n = len(dataset)
    test_size = int(n * test_ratio)
    indices = torch.randperm(n, generator=torch.Generator().manual_seed(seed))
    return data.Subset(dataset, indices[:test_size]), data.Subset(dataset, indices[test_size:])


def split_dataset(dataset, split: float, seed: Optional[int] = None):
    """reference :727-745 (torch random_split; global RNG unless a seed is given)."""
    n_train = int(len(dataset) * split)
    gen = torch.Generator().manual_seed(seed) if seed is not None else None
    kw = {} if gen is None else {'generator': gen}
    return data.random_split(dataset, [n_train, len(dataset) - n_train], **kw)


def get_dataloader(dataset, batch_size: int, shuffle: bool = True, device='cuda', drop_last: bool = True,
                   custom_collate_fn=None) -> GPUBatchLoader:
    return GPUBatchLoader(dataset, batch_size, shuffle=shuffle, drop_last=drop_last)


def get_device() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("diffgfdn_b200 needs a CUDA device")
    return torch.device('cuda')


def load_dataset(room_data: Union[RoomDataset, RIRData], device, train_valid_split_ratio: float = 0.8,
                 batch_size: int = 32, shuffle: bool = True, new_sampling_radius: Optional[float] = None,
                 drop_last: bool = False, hold_out_test_set: bool = False, test_set_ratio: Optional[float] = None,
                 test_set_seed: Optional[int] = None):
    """reference :780-867: (train, valid[, test]) loaders for a receiver grid, one loader for a single RIR."""
    if isinstance(room_data, RoomDataset):
        dataset = MultiRIRDataset(device, room_data, new_sampling_radius=new_sampling_radius)
        test_loader = None
        if hold_out_test_set:
            test_set, remaining = create_fixed_test_split(dataset, test_ratio=test_set_ratio, seed=test_set_seed)
            train_set, valid_set = split_dataset(remaining, split=train_valid_split_ratio)
            test_loader = get_dataloader(test_set, batch_size, shuffle=False, device=device, drop_last=False)
        else:
            train_set, valid_set = split_dataset(dataset, split=train_valid_split_ratio)
        train_loader = get_dataloader(train_set, batch_size, shuffle=shuffle, device=device, drop_last=drop_last)
        valid_loader = get_dataloader(valid_set, batch_size, shuffle=shuffle, device=device, drop_last=drop_last)
        return (train_loader, valid_loader, test_loader) if hold_out_test_set else (train_loader, valid_loader)
    if isinstance(room_data, RIRData):
        dataset = SingleRIRDataset(device, room_data, new_sampling_radius=new_sampling_radius)
        return get_dataloader(dataset, batch_size, shuffle=shuffle, device=device, drop_last=drop_last)
    raise TypeError("load_dataset: room_data must be a RoomDataset or RIRData")
