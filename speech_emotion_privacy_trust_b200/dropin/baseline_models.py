"""Drop-in for model/baseline_models.py: the classifiers reachable from the cloak path, with the reference's constructor
signatures and parameter names (checkpoints are loaded strictly by key, training_cloak_with_grl.py:395).  Compute stays
stock cuDNN / cuBLAS through torch.nn -- the hot path this package rewrites is extraction, cloak and gradient reversal.

    two_d_cnn_lstm            reference :143-260   3 x (conv5x5, BN, ReLU, pool2, dropout2d) -> GRU -> mean | self-attention
    deep_two_d_cnn_lstm       reference :264-385   + 4th conv block, flattened GRU output when att is None
    deep_two_d_cnn_lstm_tmp   reference :388-509   the deep model with rnn_cell defaulting to 'lstm'
    one_d_cnn_lstm            reference :19-140    3 x (conv1d k5, ReLU, pool 2/5/5) over time -> flatten | 8-head attention
    two_d_cnn                 reference :512-596   six 3x3 convs + learned time pooling (w1/w2)

The last three are not reachable from the cloak path; they exist because every reference driver imports them by name
(training_cloak_with_grl.py:24, training_cloak.py:24, training_adversary_baselines.py:24) and this module shadows the
reference's.  They keep constructor signatures, state_dict keys and forward behaviour -- including two_d_cnn's channel
mismatch (:548 -> :552), which makes its forward raise exactly as the reference's does.
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


class deep_two_d_cnn_lstm_tmp(_CnnRnnClassifier):
    deep = True

    def __init__(self, input_channel, input_spec_size, cnn_filter_size, lstm_hidden_size=128, num_layers_lstm=2,
                 pred='emotion', bidirectional=True, rnn_cell='lstm', attention_size=256, variable_lengths=False,
                 global_feature=1, att=None):
        super().__init__(input_channel, input_spec_size, cnn_filter_size, lstm_hidden_size, num_layers_lstm, pred,
                         bidirectional, rnn_cell, attention_size, variable_lengths, global_feature, att)


class one_d_cnn_lstm(nn.Module):
    """Conv1d over time with the mel bands as channels; the recurrent layer is constructed (its weights are part of the
    checkpoint) but the reference's forward never calls it (:103)."""
    N_HEADS = 8

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
        self.rnn_input_size = 512
        self.attention_size = attention_size
        self.pred = pred
        self.att = att

        cells = {'lstm': nn.LSTM, 'gru': nn.GRU}
        if rnn_cell.lower() not in cells:
            raise ValueError("Unsupported RNN Cell: {0}".format(rnn_cell))
        self.rnn_cell = cells[rnn_cell.lower()]

        self.dropout = nn.Dropout(p=self.dropout_p)
        layers = []
        for c_in, c_out, pool in ((input_spec_size, 128, 2), (128, 256, 5), (256, 512, 5)):
            layers += [nn.Conv1d(c_in, c_out, kernel_size=5, padding=2), nn.ReLU(), nn.MaxPool1d(kernel_size=pool, stride=pool),
                       nn.Dropout(self.dropout_p)]
        self.conv = nn.Sequential(*layers)
        self.rnn = self.rnn_cell(input_size=self.rnn_input_size, hidden_size=lstm_hidden_size, num_layers=num_layers_lstm,
                                 batch_first=True, dropout=self.dropout_p, bidirectional=bidirectional)
        width = lstm_hidden_size * 2
        self.att_linear1 = nn.Linear(width, attention_size)
        self.att_pool = nn.Tanh()
        self.att_linear2 = nn.Linear(attention_size, self.N_HEADS)
        self.att_mat1 = nn.Parameter(torch.rand(attention_size, width), requires_grad=True)
        self.att_mat2 = nn.Parameter(torch.rand(self.N_HEADS, attention_size), requires_grad=True)
        self.dense_relu1 = nn.ReLU()
        self.dense_relu2 = nn.ReLU()
        self.classifier = nn.Sequential(nn.Linear(512 * 4, 128), nn.ReLU(), nn.Dropout(self.dropout_p))
        self.dense2 = nn.Linear(128, 64)
        self.dense1 = nn.Linear(width + 88, 128) if global_feature == 1 else nn.Linear(512 * 4, 128)
        self.pred_emotion_layer = nn.Linear(128, self.num_emo_classes)
        self.pred_gender_layer = nn.Linear(128, self.num_gender_class)

    def forward(self, input_var, global_feature=None):
        x = self.conv(input_var.squeeze(dim=1).permute(0, 2, 1).float()).permute(0, 2, 1)      # (B, T/50, 512)
        if self.att is None:
            z = x.reshape(x.shape[0], -1)
        elif self.att == 'self_att':
            z = self_attention_pool(self, x)
        if global_feature is not None:
            z = torch.cat((z, global_feature), 1)
        z = self.classifier(z)
        if self.pred == 'multitask':
            return self.pred_emotion_layer(z), self.pred_gender_layer(z)
        return self.pred_emotion_layer(z) if self.pred == 'emotion' else self.pred_gender_layer(z)


class two_d_cnn(nn.Module):
    def __init__(self, input_channel, input_spec_size, cnn_filter_size, pred='emotion', global_feature=1, att=None):
        super().__init__()
        self.input_channel = input_channel
        self.input_spec_size = input_spec_size
        self.dropout_p = 0.5
        self.num_emo_classes = 4
        self.num_gender_class = 2
        self.cnn_filter_size = cnn_filter_size
        self.pred = pred
        self.rnn_input_size = int(64 * input_spec_size / 8)
        self.dropout = nn.Dropout(p=self.dropout_p)

        def conv(c_in, c_out):
            return nn.Conv2d(c_in, c_out, kernel_size=(3, 3), padding=(1, 1))

        def pool():
            return nn.MaxPool2d(kernel_size=(2, 2), stride=(2, 2))
        p = self.dropout_p
        self.conv = nn.Sequential(
            conv(1, 32), nn.ReLU(), nn.Dropout2d(p),
            conv(32, 48), pool(), nn.BatchNorm2d(48), nn.ReLU(), nn.Dropout2d(p),
            conv(48, 64), nn.ReLU(), nn.Dropout2d(p),
            conv(64, 64), nn.BatchNorm2d(64), nn.ReLU(), pool(), nn.Dropout2d(p),
            conv(64, 32), nn.ReLU(), nn.Dropout2d(p),
            conv(64, 64), nn.BatchNorm2d(64), nn.ReLU(), pool(), nn.Dropout2d(p),      # 32 -> 64 mismatch kept (:548/:552)
        )
        self.dense_relu1 = nn.ReLU()
        self.dense_relu2 = nn.ReLU()
        self.dense2 = nn.Linear(128, 64)
        self.dense1 = nn.Linear(2 + 88, 128) if global_feature == 1 else nn.Linear(2, 128)
        self.pred_emotion_layer = nn.Linear(128, self.num_emo_classes)
        self.pred_gender_layer = nn.Linear(128, self.num_gender_class)
        self.w1 = nn.Parameter(torch.rand(50, 4), requires_grad=True)
        self.w2 = nn.Parameter(torch.rand(50, 2), requires_grad=True)

    def forward(self, input_var, global_feature=None):
        x = self.conv(input_var.float()).transpose(1, 2).contiguous()
        x = x.reshape(x.shape[0], x.shape[1], -1).transpose(1, 2).contiguous()
        return torch.matmul(x, self.w1 if self.pred == 'emotion' else self.w2).mean(dim=1)
