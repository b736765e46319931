from layers.activations import Activation, ReLU, Softmax
from layers.attentions import MultiHeadAttention
from layers.conv import Conv2D
from layers.layer import Layer
from layers.mlp import Dense, Linear
from layers.normalizations import LayerNormalization
from layers.transformer import TransformerDecoder, TransformerEncoder
