"""MLP factory -- mirrors openrec/tf2/modules/multi_layer_perceptron.py:5-18."""
from ...tfshim.keras import Sequential
from ...tfshim.keras.layers import Dense


def MLP(units_list, use_bias=True, activation="relu", out_activation=None):
    """Sequential of Dense(units, activation) with ``out_activation`` on the last layer."""
    mlp = Sequential()
    for k, units in enumerate(units_list):
        last = k == len(units_list) - 1
        mlp.add(Dense(units, activation=out_activation if last else activation, use_bias=use_bias))
    return mlp
