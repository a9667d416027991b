defmodule Vettore.B200.Nifs do
  @moduledoc false
  # Function stubs of the erl_nif shim `nif/vettore_b200_nif.c` (ERL_NIF_INIT name
  # `Elixir.Vettore.B200.Nifs`), which maps them onto `libvettore_b200.so` (include/vettore_b200.h).
  #
  # The first block carries the scan-path functions of `Vettore.Nifs` with the same names, arities,
  # argument order and result shapes (reference lib/vettore_nifs.ex:70-172), so a caller can alias this
  # module where it aliased `Vettore.Nifs`. HNSW, MUVERA, the pairwise metrics and the normalisers are
  # not here: they stay in the Rust library. The second block is additive (resident pipelines).
  #
  # tests/test_nif_shim.py asserts that the stubs below and the shim's ErlNifFunc table are the same set.

  @on_load :load_nif

  @doc false
  def load_nif do
    path = :filename.join(:code.priv_dir(:vettore), ~c"native/libvettore_b200_nif")
    :erlang.load_nif(path, 0)
  end

  # ---- scan-path subset of Vettore.Nifs ------------------------------------------------------------
  @spec compress_sign_bits([float()]) :: [non_neg_integer()]
  def compress_sign_bits(_vector), do: :erlang.nif_error(:nif_not_loaded)

  @spec vector_top_k([{String.t(), [float()]}], [float()], non_neg_integer(), pos_integer(), non_neg_integer()) ::
          {:ok, [{String.t(), float()}]} | {:error, String.t()}
  def vector_top_k(_vectors, _query, _metric_code, _dimensions, _limit), do: :erlang.nif_error(:nif_not_loaded)

  @spec binary_top_k([{String.t(), [non_neg_integer()]}], [non_neg_integer()], pos_integer(), non_neg_integer()) ::
          {:ok, [{String.t(), float()}]} | {:error, String.t()}
  def binary_top_k(_vectors, _query, _dimensions, _limit), do: :erlang.nif_error(:nif_not_loaded)

  @spec multi_vector_score([[float()]], [[float()]], non_neg_integer()) :: {:ok, float()} | {:error, String.t()}
  def multi_vector_score(_query_vectors, _document_vectors, _metric_code), do: :erlang.nif_error(:nif_not_loaded)

  @spec multi_vector_top_k([{String.t(), [[float()]]}], [[float()]], non_neg_integer(), non_neg_integer()) ::
          {:ok, [{String.t(), float()}]} | {:error, String.t()}
  def multi_vector_top_k(_documents, _query_vectors, _metric_code, _limit), do: :erlang.nif_error(:nif_not_loaded)

  @spec flat_new_l2() :: reference()
  def flat_new_l2, do: :erlang.nif_error(:nif_not_loaded)
  @spec flat_new_l2_squared() :: reference()
  def flat_new_l2_squared, do: :erlang.nif_error(:nif_not_loaded)
  @spec flat_new_cosine() :: reference()
  def flat_new_cosine, do: :erlang.nif_error(:nif_not_loaded)
  @spec flat_new_inner_product() :: reference()
  def flat_new_inner_product, do: :erlang.nif_error(:nif_not_loaded)
  @spec flat_new_negative_inner_product() :: reference()
  def flat_new_negative_inner_product, do: :erlang.nif_error(:nif_not_loaded)
  @spec flat_new_manhattan() :: reference()
  def flat_new_manhattan, do: :erlang.nif_error(:nif_not_loaded)
  @spec flat_new_chebyshev() :: reference()
  def flat_new_chebyshev, do: :erlang.nif_error(:nif_not_loaded)
  @spec flat_new_hamming() :: reference()
  def flat_new_hamming, do: :erlang.nif_error(:nif_not_loaded)
  @spec flat_new_jaccard() :: reference()
  def flat_new_jaccard, do: :erlang.nif_error(:nif_not_loaded)

  @spec flat_insert(reference(), String.t(), [float()]) :: {:ok, {}} | {:error, String.t()}
  def flat_insert(_index, _id, _vector), do: :erlang.nif_error(:nif_not_loaded)

  @spec flat_insert_many(reference(), [{String.t(), [float()]}]) :: {:ok, {}} | {:error, String.t()}
  def flat_insert_many(_index, _vectors), do: :erlang.nif_error(:nif_not_loaded)

  @spec flat_delete(reference(), String.t()) :: {:ok, {}} | {:error, String.t()}
  def flat_delete(_index, _id), do: :erlang.nif_error(:nif_not_loaded)

  @spec flat_search(reference(), [float()], pos_integer()) :: {:ok, [{String.t(), float()}]} | {:error, String.t()}
  def flat_search(_index, _query, _limit), do: :erlang.nif_error(:nif_not_loaded)

  @spec muvera_encode_query([[float()]], pos_integer(), pos_integer(), non_neg_integer(), non_neg_integer(), pos_integer(), pos_integer() | nil) ::
          {:ok, [float()]} | {:error, String.t()}
  def muvera_encode_query(_vectors, _dimension, _num_repetitions, _num_simhash_projections, _seed, _projection_dimension, _final_projection_dimension), do: :erlang.nif_error(:nif_not_loaded)

  @spec muvera_encode_document([[float()]], pos_integer(), pos_integer(), non_neg_integer(), non_neg_integer(), pos_integer(), pos_integer() | nil) ::
          {:ok, [float()]} | {:error, String.t()}
  def muvera_encode_document(_vectors, _dimension, _num_repetitions, _num_simhash_projections, _seed, _projection_dimension, _final_projection_dimension), do: :erlang.nif_error(:nif_not_loaded)

  # ---- additive: resident pipelines (no counterpart in Vettore.Nifs) -------------------------------
  # Capacity hint before a snapshot is streamed in (Collection.rebuild_index).
  @spec flat_reserve(reference(), non_neg_integer()) :: {:ok, {}} | {:error, String.t()}
  def flat_reserve(_index, _rows), do: :erlang.nif_error(:nif_not_loaded)

  # flat_search plus Distance.result_values/3 for every hit: [{id, raw, score, distance}].
  # score_mode: 0 = :raw, 1 = :similarity.
  @spec flat_search_shaped(reference(), [float()], pos_integer(), non_neg_integer(), 0 | 1) ::
          {:ok, [{String.t(), float(), float(), float()}]} | {:error, String.t()}
  def flat_search_shaped(_index, _query, _limit, _metric_code, _score_mode), do: :erlang.nif_error(:nif_not_loaded)

  # funnel_search on the resident matrix: every stage and the exact rerank in one call.
  @spec flat_funnel_search(reference(), [float()], non_neg_integer(), [pos_integer()], pos_integer(), pos_integer()) ::
          {:ok, [{String.t(), float()}]} | {:error, String.t()}
  def flat_funnel_search(_index, _query, _metric_code, _stages, _candidates, _limit), do: :erlang.nif_error(:nif_not_loaded)

  # quantized_search on the resident sign codes + exact rerank in one call.
  @spec flat_quantized_search(reference(), [float()], non_neg_integer(), pos_integer(), pos_integer()) ::
          {:ok, [{String.t(), float()}]} | {:error, String.t()}
  def flat_quantized_search(_index, _query, _metric_code, _candidates, _limit), do: :erlang.nif_error(:nif_not_loaded)

  # HBM-resident multi-vector collection (multi_vector_search without store.all + by-value marshalling).
  @spec mv_new(non_neg_integer()) :: {:ok, reference()} | {:error, String.t()}
  def mv_new(_metric_code), do: :erlang.nif_error(:nif_not_loaded)

  @spec mv_insert_many(reference(), [{String.t(), [[float()]]}]) :: {:ok, {}} | {:error, String.t()}
  def mv_insert_many(_index, _documents), do: :erlang.nif_error(:nif_not_loaded)

  @spec mv_delete(reference(), String.t()) :: {:ok, {}} | {:error, String.t()}
  def mv_delete(_index, _id), do: :erlang.nif_error(:nif_not_loaded)

  @spec mv_search(reference(), [[float()]], pos_integer()) :: {:ok, [{String.t(), float()}]} | {:error, String.t()}
  def mv_search(_index, _query_vectors, _limit), do: :erlang.nif_error(:nif_not_loaded)
end
