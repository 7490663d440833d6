"""Drop-in for model/baseline_models.py: the classifiers reachable from the cloak path, with the reference's constructor
signatures and parameter names (checkpoints are loaded strictly by key, training_cloak_with_grl.py:395).  Compute stays
stock cuDNN / cuBLAS through torch.nn -- the hot path this package rewrites is extraction, cloak and gradient reversal.

    two_d_cnn_lstm        reference :143-260   3 x (conv5x5, BN, ReLU, pool2, dropout2d) -> GRU -> mean | self-attention
    deep_two_d_cnn_lstm   reference :264-385   + 4th conv block, flattened GRU output when att is None
"""
import torch
import torch.nn as nn

N_ATT_HEADS = 16


def _conv_block(c_in, c_out, p_drop, pool=True):
    layers = [nn.Conv2d(c_in, c_out, kernel_size=5, padding=2), nn.BatchNorm2d(c_out), nn.ReLU()]
    if pool:
        layers.append(nn.MaxPool2d(kernel_size=2, stride=2))
    layers.append(nn.Dropout2d(p_drop))
    return layers


def sequence_features(model, x):
    """conv stack -> (B, T', C*F') -> recurrent layers; shared by the plain classifiers and the cloak wrappers."""
    x = model.conv(x.float())
    x = x.transpose(1, 2).contiguous()
    b, t = x.shape[0], x.shape[1]
    x, _ = model.rnn(x.reshape(b, t, -1))
    return x


def self_attention_pool(model, x):
    att = model.att_linear2(model.att_pool(model.att_linear1(x))).transpose(1, 2)
    return torch.matmul(torch.softmax(att, dim=2), x).mean(dim=1)


def classify(model, z, global_feature=None, pred=None):
    if global_feature is not None:
        z = torch.cat((z, global_feature), 1)
    z = model.dropout(model.dense_relu1(model.dense1(z)))
    pred = model.pred if pred is None else pred
    if pred == 'multitask':
        return model.pred_emotion_layer(z), model.pred_gender_layer(z)
    if pred == 'emotion':
        return model.pred_emotion_layer(z)
    return model.pred_gender_layer(z)


class _CnnRnnClassifier(nn.Module):
    deep = False

    def __init__(self, input_channel, input_spec_size, cnn_filter_size, lstm_hidden_size=128, num_layers_lstm=2,
                 pred='emotion', bidirectional=True, rnn_cell='gru', attention_size=256, variable_lengths=False,
                 global_feature=1, att=None):
        super().__init__()
        self.input_channel = input_channel
        self.input_spec_size = input_spec_size
        self.lstm_hidden_size = lstm_hidden_size
        self.bidirectional = bidirectional
        self.num_layers_lstm = num_layers_lstm
        self.dropout_p = 0.2
        self.variable_lengths = variable_lengths
        self.num_emo_classes = 4
        self.num_gender_class = 2
        self.cnn_filter_size = cnn_filter_size
        self.attention_size = attention_size
        self.pred = pred
        self.att = att
        self.rnn_input_size = int(128 * input_spec_size / 8)

        cells = {'lstm': nn.LSTM, 'gru': nn.GRU}
        if rnn_cell.lower() not in cells:
            raise ValueError("Unsupported RNN Cell: {0}".format(rnn_cell))
        self.rnn_cell = cells[rnn_cell.lower()]

        self.dropout = nn.Dropout(p=self.dropout_p)
        blocks = _conv_block(1, 32, self.dropout_p) + _conv_block(32, 64, self.dropout_p) + _conv_block(64, 128, self.dropout_p)
        if self.deep:
            blocks += _conv_block(128, 128, self.dropout_p, pool=False)
        self.conv = nn.Sequential(*blocks)
        self.rnn = self.rnn_cell(input_size=self.rnn_input_size, hidden_size=lstm_hidden_size, num_layers=num_layers_lstm,
                                 batch_first=True, dropout=self.dropout_p, bidirectional=bidirectional)

        width = lstm_hidden_size * 2
        self.att_linear1 = nn.Linear(width, attention_size, bias=False)
        self.att_pool = nn.Tanh()
        self.att_linear2 = nn.Linear(attention_size, N_ATT_HEADS, bias=False)
        self.att_mat1 = nn.Parameter(torch.rand(attention_size, width), requires_grad=True)
        self.att_mat2 = nn.Parameter(torch.rand(N_ATT_HEADS, attention_size), requires_grad=True)
        self.dense_relu1 = nn.ReLU()
        self.dense_relu2 = nn.ReLU()
        self.dense2 = nn.Linear(128, 64)
        if global_feature == 1:
            self.dense1 = nn.Linear(width + 88, 128)
        else:
            self.dense1 = nn.Linear(width * 25 if self.deep else width, 128)
        self.pred_emotion_layer = nn.Linear(128, self.num_emo_classes)
        self.pred_gender_layer = nn.Linear(128, self.num_gender_class)
        # the reference's init_weight() walks module NAMES and therefore changes nothing (:213-220): default init stays

    def pool(self, x):
        if self.att == 'self_att':
            return self_attention_pool(self, x)
        if self.deep:
            return x.reshape(x.shape[0], -1)
        return x.mean(dim=1)

    def forward(self, input_var, global_feature=None):
        return classify(self, self.pool(sequence_features(self, input_var)), global_feature)


class two_d_cnn_lstm(_CnnRnnClassifier):
    deep = False


class deep_two_d_cnn_lstm(_CnnRnnClassifier):
    deep = True
