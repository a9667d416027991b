defmodule Vettore.Index.B200 do
  @moduledoc """
  Exact flat index resident in the HBM of a B200 (`index: Vettore.Index.B200`).

  Same contract as the stock flat index: ETS stays the canonical store of values and metadata, the
  device mirrors ids and vectors, `search/3` is one native call followed by hydration from the store
  (ids the store no longer has are dropped). What differs from `Vettore.Index.Flat`:

    * the native calls go to `Vettore.B200.Nifs` (CUDA, `libvettore_b200.so`) instead of the Rust scan;
    * result shaping is batched: `flat_search_shaped/5` returns `{id, raw, score, distance}` for the
      whole hit list (`Distance.result_values/3` evaluated natively), so hydration is the store lookups
      and nothing else;
    * `funnel_search/4` and `quantized_search/4` run those pipelines on the resident matrix / sign codes
      instead of `store.all` + by-value NIFs; `Vettore.Collection` may call them when the index module
      exports them (public API unchanged).

  This module cannot be compiled in the build image (no Elixir/OTP there); it is the reference-side half
  of the drop-in described in INTEGRATION.md and is kept in step with the shim by
  `tests/test_nif_shim.py`.
  """

  @behaviour Vettore.Index

  alias Vettore.B200.Nifs
  alias Vettore.{Collection, Embedding, Result}

  @usize_max 4_294_967_295
  @metric_codes %{
    l2: 0,
    l2_squared: 1,
    cosine: 2,
    inner_product: 3,
    negative_inner_product: 4,
    manhattan: 5,
    chebyshev: 6,
    hamming: 7,
    jaccard: 8
  }

  # ---- Vettore.Index callbacks ---------------------------------------------------------------------
  @impl true
  def new(metric, opts \\ [])

  def new(metric, []) when is_map_key(@metric_codes, metric) do
    {:ok, apply(Nifs, :"flat_new_#{metric}", [])}
  rescue
    # the constructors raise when no CUDA device is present: surface it as an index error
    error -> {:error, {:b200_unavailable, Exception.message(error)}}
  end

  def new(metric, []), do: {:error, {:unsupported_flat_metric, metric}}
  def new(_metric, _opts), do: {:error, :invalid_flat_options}

  @impl true
  def put(%Collection{index_state: index}, %Embedding{id: id, vector: vector}) do
    unit(Nifs.flat_insert(index, id, vector))
  end

  @impl true
  def put_many(%Collection{index_state: index}, embeddings) when is_list(embeddings) do
    # one native call, one bulk host->device copy; ids arrive sorted from rebuild_index, which is the
    # O(1)-per-row append path of the device-side id table
    if length(embeddings) > 4096, do: Nifs.flat_reserve(index, length(embeddings))
    unit(Nifs.flat_insert_many(index, for(%Embedding{id: id, vector: v} <- embeddings, do: {id, v})))
  end

  @impl true
  def delete(%Collection{index_state: index}, id), do: unit(Nifs.flat_delete(index, id))

  @impl true
  def search(%Collection{} = collection, query, opts) do
    with {:ok, limit} <- limit_option(opts),
         {:ok, query} <- Collection.prepare_query(collection, query),
         {:ok, shaped} <-
           Nifs.flat_search_shaped(
             collection.index_state,
             query,
             limit,
             Map.fetch!(@metric_codes, collection.metric),
             score_mode(collection.score)
           ) do
      {:ok, hydrate(collection, shaped)}
    end
  end

  # ---- additive: resident pipelines ------------------------------------------------------------------
  @doc "funnel_search on the resident matrix: `stages` prefix lengths, `candidates` survivors per stage."
  def funnel_search(%Collection{} = collection, query, stages, opts) when is_list(stages) do
    with {:ok, limit} <- limit_option(opts),
         candidates = Keyword.get(opts, :candidates, max(limit * 10, limit)),
         {:ok, query} <- Collection.prepare_query(collection, query),
         {:ok, hits} <-
           Nifs.flat_funnel_search(collection.index_state, query, Map.fetch!(@metric_codes, collection.metric), stages, candidates, limit) do
      {:ok, hydrate_raw(collection, hits)}
    end
  end

  @doc "quantized_search on the resident sign codes, exact rerank of the candidates on the device."
  def quantized_search(%Collection{} = collection, query, opts) do
    with {:ok, limit} <- limit_option(opts),
         candidates = Keyword.get(opts, :candidates, max(limit * 10, limit)),
         {:ok, query} <- Collection.prepare_query(collection, query),
         {:ok, hits} <-
           Nifs.flat_quantized_search(collection.index_state, query, Map.fetch!(@metric_codes, collection.metric), candidates, limit) do
      {:ok, hydrate_raw(collection, hits)}
    end
  end

  # ---- helpers ---------------------------------------------------------------------------------------
  defp unit({:ok, {}}), do: :ok
  defp unit(:ok), do: :ok
  defp unit({:error, _} = error), do: error

  defp score_mode(:similarity), do: 1
  defp score_mode(_raw), do: 0

  defp limit_option(opts) when is_list(opts) do
    cond do
      not Keyword.keyword?(opts) -> {:error, :invalid_search_options}
      Keyword.keys(opts) -- [:limit, :candidates] != [] -> {:error, :invalid_search_options}
      true -> check_limit(Keyword.get(opts, :limit, 10))
    end
  end

  defp limit_option(_opts), do: {:error, :invalid_search_options}

  defp check_limit(limit) when is_integer(limit) and limit in 1..@usize_max, do: {:ok, limit}
  defp check_limit(_limit), do: {:error, :invalid_limit}

  # Hits already carry score and distance: only the store lookups remain. Ids gone from the store
  # (a delete that raced the search, or phantom ids after a failed store write) are dropped.
  defp hydrate(collection, shaped) do
    for {id, _raw, score, distance} <- shaped,
        {:ok, %Embedding{} = embedding} <- [Collection.get(collection, id)] do
      %Result{id: id, value: embedding.value, score: score, distance: distance, metric: collection.metric, metadata: embedding.metadata}
    end
  end

  defp hydrate_raw(collection, hits) do
    for {id, raw} <- hits,
        {:ok, %Embedding{} = embedding} <- [Collection.get(collection, id)] do
      {score, distance} = Vettore.Distance.result_values(collection.metric, raw, collection.score)
      %Result{id: id, value: embedding.value, score: score, distance: distance, metric: collection.metric, metadata: embedding.metadata}
    end
  end
end
