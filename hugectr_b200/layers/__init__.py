"""Layer registry: Layer_t -> implementation (the role of add_dense_layer_impl's giant switch,
HugeCTR/src/pybind/add_dense_layer_helpers.cpp:122-877)."""
from ..enums import Layer_t
from .base import BuildCtx, Layer, Param, ParamArena, TensorBag, TorchLayer
from .common import (AddLayer, BatchNormLayer, CastLayer, ConcatLayer, DropoutLayer, ELULayer,
                     ElementwiseMultiplyLayer, FmOrder2Layer, FusedReshapeConcatGeneralLayer,
                     FusedReshapeConcatLayer, GatherLayer, GRULayer, LayerNormLayer,
                     MatrixMultiplyLayer, MultiHeadAttentionLayer, PReLUDiceLayer, ReduceMeanLayer,
                     ReduceSumLayer, ReLULayer, ReshapeLayer, ScaleLayer, SelectLayer,
                     SequenceMaskLayer, SigmoidLayer, SliceLayer, SoftmaxLayer, SubLayer,
                     WeightMultiplyLayer)
from .cross import InteractionLayer, MultiCrossLayer
from .loss import (BinaryCrossEntropyLossLayer, CrossEntropyLossLayer, MultiCrossEntropyLossLayer,
                   Regularizer)
from .mlp import FusedInnerProductLayer, InnerProductLayer, MLPLayer

LAYER_REGISTRY = {
    Layer_t.BatchNorm: BatchNormLayer,
    Layer_t.LayerNorm: LayerNormLayer,
    Layer_t.BinaryCrossEntropyLoss: BinaryCrossEntropyLossLayer,
    Layer_t.Reshape: ReshapeLayer,
    Layer_t.Select: SelectLayer,
    Layer_t.Concat: ConcatLayer,
    Layer_t.Concat3D: ConcatLayer,
    Layer_t.CrossEntropyLoss: CrossEntropyLossLayer,
    Layer_t.Dropout: DropoutLayer,
    Layer_t.ElementwiseMultiply: ElementwiseMultiplyLayer,
    Layer_t.DotProduct: ElementwiseMultiplyLayer,
    Layer_t.ELU: ELULayer,
    Layer_t.InnerProduct: InnerProductLayer,
    Layer_t.FusedInnerProduct: FusedInnerProductLayer,
    Layer_t.MLP: MLPLayer,
    Layer_t.Interaction: InteractionLayer,
    Layer_t.MultiCrossEntropyLoss: MultiCrossEntropyLossLayer,
    Layer_t.ReLU: ReLULayer,
    Layer_t.ReLUHalf: ReLULayer,
    Layer_t.Sigmoid: SigmoidLayer,
    Layer_t.Slice: SliceLayer,
    Layer_t.WeightMultiply: WeightMultiplyLayer,
    Layer_t.FmOrder2: FmOrder2Layer,
    Layer_t.Add: AddLayer,
    Layer_t.ReduceSum: ReduceSumLayer,
    Layer_t.Softmax: SoftmaxLayer,
    Layer_t.MaskedSoftmax: SoftmaxLayer,
    Layer_t.Gather: GatherLayer,
    Layer_t.PReLU_Dice: PReLUDiceLayer,
    Layer_t.GRU: GRULayer,
    Layer_t.MatrixMultiply: MatrixMultiplyLayer,
    Layer_t.MultiHeadAttention: MultiHeadAttentionLayer,
    Layer_t.Scale: ScaleLayer,
    Layer_t.FusedReshapeConcat: FusedReshapeConcatLayer,
    Layer_t.FusedReshapeConcatGeneral: FusedReshapeConcatGeneralLayer,
    Layer_t.Sub: SubLayer,
    Layer_t.ReduceMean: ReduceMeanLayer,
    Layer_t.MultiCross: MultiCrossLayer,
    Layer_t.Cast: CastLayer,
    Layer_t.SequenceMask: SequenceMaskLayer,
}

TRAINABLE_LAYERS = {Layer_t.InnerProduct, Layer_t.MultiCross, Layer_t.WeightMultiply,
                    Layer_t.BatchNorm, Layer_t.LayerNorm, Layer_t.GRU, Layer_t.MultiHeadAttention,
                    Layer_t.MLP, Layer_t.FusedInnerProduct}
LOSS_LAYERS = {Layer_t.BinaryCrossEntropyLoss, Layer_t.CrossEntropyLoss,
               Layer_t.MultiCrossEntropyLoss}
