"""Drop-in for the reference's ``scripts/getEmbeddingExample.py`` (same functions and command line), with the feature
extraction, the front-end, the pooling and the FC tail all on the GPU:

    python -m doubleattentionspeakerverification_b200.getEmbeddingExample --audioPath utt.wav \\
        --modelConfig config.pkl --modelCheckpoint model.chkpt

``--device`` is accepted for compatibility; this package has no CPU path, so anything but ``cuda`` is an error.
"""
import argparse
import pickle

import torch

from .featureExtractor import extractFeatures
from .model import SpeakerClassifier


def prepareInput(features, device):
    """scripts/getEmbeddingExample.py:7-12."""
    return torch.as_tensor(features, dtype=torch.float32).to(device).unsqueeze(0)


def getAudioEmbedding(audioPath, net, device):
    """scripts/getEmbeddingExample.py:15-20."""
    features = extractFeatures(audioPath)
    with torch.no_grad():
        return net.getEmbedding(prepareInput(features, device))


def main(opt, params):
    """scripts/getEmbeddingExample.py:23-39."""
    if params.device != 'cuda':
        raise SystemExit('this package runs on CUDA only (--device cuda)')
    print('Loading Model')
    device = torch.device('cuda')
    net_dict = torch.load(params.modelCheckpoint, map_location=device, weights_only=False)
    opt = net_dict['settings']
    print(torch.cuda.get_device_name(0))
    net = SpeakerClassifier(opt, device)
    net.load_state_dict(net_dict['model'])
    net.to(device)
    net.eval()
    embedding = getAudioEmbedding(params.audioPath, net, device)
    print(embedding)
    return embedding


if __name__ == '__main__':
    parser = argparse.ArgumentParser(description='score a trained model')
    parser.add_argument('--audioPath', type=str, required=True)
    parser.add_argument('--modelConfig', type=str, required=True)
    parser.add_argument('--modelCheckpoint', type=str, required=True)
    parser.add_argument('--device', type=str, default='cuda', choices=['cpu', 'cuda'])
    params = parser.parse_args()
    with open(params.modelConfig, 'rb') as handle:
        opt = pickle.load(handle)
    main(opt, params)
